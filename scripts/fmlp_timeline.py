"""clock64 timeline of the CTA-pair fused MLP kernel (cluster 0 leader, its second tile) + per-layer kernel times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from vitmarl_b200 import vit, _capi
lib = _capi.lib()
cfg = vit.ViTConfig(64, 64, 2, 8, 192, 1, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
x = (torch.rand(4096, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(2): enc.apply_packed(packed, x)
enc.options.debug_timeline = buf.data_ptr()
if len(sys.argv) > 1: enc.options.attn_flags = int(sys.argv[1], 0) & 0xff
enc.apply_packed(packed, x); torch.cuda.synchronize()
enc.options.debug_timeline = None
t = buf.cpu().tolist(); t0 = t[0]
r = lambda i: t[i] - t0 if t[i] else None
print("compute: tile_start 0  (stamps relative, cycles)")
for c in range(6):
    print(f" c{c}: acc1full {r(10+4*c)} gelu_done {r(11+4*c)} hready {r(12+4*c)} | mma: fc1 start {r(100+4*c)} issued {r(101+4*c)} fc2 start {r(102+4*c)} issued {r(103+4*c)}")
print("LN(next) done", r(50), "acc2full", r(60), "tile end", r(61))
print("issuer wait cycles (tile 1): xnready", t[200], "acc1empty", t[201], "w1full", t[202], "acc2empty", t[203], "hready", t[204], "w2full", t[205])
print("compute warp 4 wait cycles: acc1full", t[210], "hempty", t[211], "acc2full", t[212])
# per-category kernel time over 5 forwards
tm = _capi.Timing(); enc.options.timing = tm.handle
for _ in range(5): enc.apply_packed(packed, x)
torch.cuda.synchronize()
ms, n, _ = tm.read(); enc.options.timing = None
names = ["gemm", "fused_mlp", "fused_attn", "attention", "layernorm", "other", "dw", "dx"]
print({nm: round(ms[i] / max(n[i], 1) * 1e3, 1) for i, nm in enumerate(names) if n[i]}, "us per launch")
