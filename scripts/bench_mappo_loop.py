"""BASELINE configs[4] shape: full rollout + update iteration (see vitmarl_b200/mappo_loop.py).

    python scripts/bench_mappo_loop.py [--envs 8192 --rollout 128 --epochs 4 --minibatches 16 --micro 8192]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_mappo_loop.py
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import mappo_loop

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=8192)          # per GPU (65536 over 8 GPUs)
ap.add_argument("--rollout", type=int, default=128)        # NUM_STEPS
ap.add_argument("--epochs", type=int, default=4)           # UPDATE_EPOCHS
ap.add_argument("--minibatches", type=int, default=16)     # NUM_MINIBATCHES
ap.add_argument("--micro", type=int, default=8192)         # images per forward+backward launch sequence
ap.add_argument("--msgs", type=int, default=13)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
loop = mappo_loop.MappoLoop(a.envs, a.rollout, a.epochs, a.minibatches, a.micro, a.msgs, rank=rank, world=world)
res = loop.run(dist)
if rank == 0:
    print(json.dumps(res))
if dist is not None:
    dist.destroy_process_group()
