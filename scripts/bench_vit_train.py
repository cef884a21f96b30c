"""ViT fwd+bwd ms (BASELINE metric iii; configs[3]: ViT-S/16 bf16 on 128x128 LOB images, batch 8192) and, under
torchrun, the data-parallel training step with the PPO-gradient pmean over NCCL (configs[4] shape: 8192 images/GPU).

    python scripts/bench_vit_train.py [--model small16|tiny8] [--batch 8192] [--steps 5]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_vit_train.py ...
"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi, parallel, vit

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small16")
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = vit.VIT_SMALL_16 if a.model == "small16" else vit.VIT_TINY_8
lib = _capi.lib()
params = vit.init_params(cfg, 0, "cuda")
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, params)
B = a.batch
g = torch.Generator(device="cpu").manual_seed(rank)
lens = torch.randint(0, cfg.img_w + 1, (B, cfg.img_h, 1, cfg.channels), generator=g)
x = (torch.arange(cfg.img_w)[None, None, :, None] < lens).to(torch.bfloat16).cuda()
dy = torch.randn(B, cfg.dim, generator=g).cuda()
red = parallel.GradAllReducer([t.shape for t in packed], device="cuda", bucket_ranges=enc.bucket_param_ranges())

def step():
    enc.apply_packed(packed, x, train=True)
    enc.vjp_packed(packed, dy, grads=red.grads(), flat=red.flat, bucket_events=red.events)
    red.allreduce_mean(async_op=True)      # per-block buckets behind the backward's events (overlaps the remaining blocks)
    red.swap()                             # the next step's backward writes the other flat buffer
    red.wait()

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

for _ in range(a.warmup): step()
sync()
tm = _capi.Timing(); enc.options.timing = tm.handle
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): step()
e1.record(); sync()
ms = e0.elapsed_time(e1) / a.steps
cat_ms, cat_n, flops = tm.read()
class _V:      # (kept field names of the old report)
    pass
gm, gf = _V(), _V()
gm.value = sum(cat_ms[i] for i in (0, 1, 2, 6, 7)); gf.value = flops
names = ["gemm_fwd", "fused_mlp", "fused_attn", "attention", "layernorm", "other", "gemm_dW", "gemm_dX"]
breakdown = {nm: round(cat_ms[i] / a.steps, 3) for i, nm in enumerate(names) if cat_n[i]}
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
T, D, L, P, C = cfg.tokens, cfg.dim, cfg.depth, cfg.patch, cfg.channels
F = 2 * T * P * P * C * D + L * (24 * T * D * D + 4 * T * T * D)
if rank == 0:
    print(json.dumps({"metric": "vit_fwd_bwd_ms", "value": ms, "unit": "ms", "model": a.model, "batch_per_gpu": B, "n_gpus": world,
                      "alg_tflop_per_step": 3 * F * B / 1e12, "achieved_tflops_per_gpu": 3 * F * B / (ms * 1e-3) / 1e12,
                      "gemm_ms_per_step": gm.value / a.steps, "gemm_tflops": gf.value / (gm.value * 1e-3) / 1e12,
                      "grad_bytes_allreduced": red.flat.numel() * 4 if world > 1 else 0,
                      "workspace_GB": enc._ws.numel() / 1e9, "ms_by_kernel_class": breakdown}))
if world > 1: dist.destroy_process_group()
