"""BASELINE configs[4] shape: a full rollout + update iteration, environments sharded over the GPUs of one box.

    rollout : S steps of  order-book step -> render -> ViT-Tiny/8 encode  on E envs per GPU (RolloutEncoder.step), the rendered
              observation of every step kept for the update (S x E x 16 KB);
    update  : `epochs` passes over the S*E stored observations in `minibatches` minibatches; a minibatch is ViT forward+backward
              over its images (micro-batches of `micro`, gradients accumulated) followed by ONE NCCL mean of the packed fp32
              gradient table (the `jax.lax.pmean(grads)` of ippo_rnn_JAXMARL_pmap.py:565) -- epochs*minibatches collectives.
The policy head, PPO loss and optimiser are boundary-only rows of SURVEY.md 8a (A13-A14): dL/d(encoding) is synthetic and the
parameters are not changed, but every encoder FLOP, every byte and every collective of the iteration is there.

    python scripts/bench_mappo_loop.py [--envs 8192 --rollout 128 --epochs 4 --minibatches 16 --micro 8192]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_mappo_loop.py
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vitmarl_b200 import jaxob, parallel, rollout, synth, vit
from vitmarl_b200.config import World_EnvironmentConfig

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

cfg, vcfg = World_EnvironmentConfig(), vit.VIT_TINY_8
E, S, M = a.envs, a.rollout, a.msgs
seed = 1234 + 1000 * rank
l2 = synth.make_l2_books(E, seed)
init = torch.from_numpy(synth.init_msgs_from_l2_batched(l2)).cuda()
asks0, bids0, _ = jaxob.scan_through_entire_array(cfg, None, init, (jaxob.init_orderside(100, E), jaxob.init_orderside(100, E), None))
stream = synth.MessageStream(E, seed)
blocks = [torch.from_numpy(stream.next(M)).cuda() for _ in range(8)]       # cycled (the books keep evolving)
params = vit.init_params(vcfg, 0, "cuda")
eng = rollout.RolloutEncoder(cfg, vcfg, params, E, M)
enc = vit.ViTEncoder(vcfg)                                                  # training path (own workspace)
packed = vit.pack_params(vcfg, params)
red = parallel.GradAllReducer([t.shape for t in packed], device="cuda")
acc = torch.zeros_like(red.flat)
traj = torch.empty((S, E, vcfg.img_h, vcfg.img_w, vcfg.channels), dtype=torch.bfloat16, device="cuda")
feats = torch.empty((S, E, vcfg.dim), dtype=torch.float32, device="cuda")
total = S * E
mb_size = total // a.minibatches
micro = min(a.micro, mb_size)
dy = torch.randn(micro, vcfg.dim, device="cuda")
flat_traj = traj.view(total, vcfg.img_h, vcfg.img_w, vcfg.channels)


def sync():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def do_rollout():
    for t in range(S):
        feats[t].copy_(eng.step(blocks[t % len(blocks)]))
        traj[t].copy_(eng.last_image())


def do_update():
    for _ in range(a.epochs):
        for mb in range(a.minibatches):
            acc.zero_()
            for lo in range(mb * mb_size, (mb + 1) * mb_size, micro):
                x = flat_traj[lo:lo + micro]
                enc.apply_packed(packed, x, train=True)
                enc.vjp_packed(packed, dy[: x.shape[0]], grads=red.grads())
                acc.add_(red.flat)
            red.flat.copy_(acc)
            red.allreduce_mean()                                            # pmean of the minibatch gradient
    eng.encoder._fold_key = None                                            # the rollout engine re-folds after an update


eng.reset(asks0.clone(), bids0.clone())
for t in range(3):
    eng.step(blocks[t])
x = flat_traj[:micro]; traj[0].copy_(eng.last_image())
enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy, grads=red.grads()); red.allreduce_mean()
sync()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record(); do_rollout(); ev[1].record(); do_update(); ev[2].record()
sync()
t_roll, t_upd = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
if dist is not None:
    tt = torch.tensor([t_roll, t_upd, t_roll + t_upd], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_roll, t_upd, t_all = [float(v) for v in tt]
else:
    t_all = t_roll + t_upd
if rank == 0:
    print(json.dumps({"metric": "mappo_iteration_env_steps_per_sec", "value": world * E * S / t_all, "unit": "env-steps/s", "n_gpus": world,
                      "envs_total": world * E, "rollout_steps": S, "epochs": a.epochs, "minibatches": a.minibatches, "micro_batch": micro,
                      "rollout_s": t_roll, "update_s": t_upd, "rollout_env_steps_per_sec": world * E * S / t_roll,
                      "update_images_per_sec": world * a.epochs * total / t_upd,
                      "collectives": a.epochs * a.minibatches if world > 1 else 0, "grad_bytes_per_collective": red.flat.numel() * 4,
                      "note": "encoder + env + NCCL only: policy head / PPO loss / optimiser are out of scope (synthetic dL/d(encoding))"}))
if dist is not None:
    dist.destroy_process_group()
