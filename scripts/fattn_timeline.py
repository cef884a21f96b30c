"""clock64 timeline of the fused attention block kernel (CTA 0, its second tile) + per-layer kernel times."""
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
t = buf.cpu().tolist()[256:]
t0 = min(v for v in t if v)
r = lambda i: t[i] - t0 if t[i] else None
for h in range(3):
    b = 4 * h
    print(f" h{h}: CV: qkvfull {r(30+2*h)} epi_done {r(31+2*h)} | SM: sfull {r(10+b)} p_done {r(11+b)} oc_done(CV) {r(13+b)} | MMA: S {r(100+b)} QKVnext {r(101+b)} PV {r(102+b)}")
print(" epi(h1): ld+pack done", r(70), "waits done", r(71), "sts issued", r(72), "fence done", r(73))
print(" LN(next): start", r(51), "xfull", r(52), "stats+bar", r(53), "xnfree", r(54))
print(" LN(next) done", r(50), "| proj start", r(130), "issued", r(131), "| CV: projfull", r(60), "tile end", r(61))
ts = [t[140 + j] for j in range(16) if t[140 + j]]
print(" S(h0) issue per tile: periods:", [b - a for a, b in zip(ts, ts[1:])])
tm = _capi.Timing(); enc.options.timing = tm.handle
for _ in range(5): enc.apply_packed(packed, x)
torch.cuda.synchronize()
ms, n, _ = tm.read(); enc.options.timing = None
names = ["gemm", "fused_mlp", "fused_attn", "attention", "layernorm", "other", "dw", "dx"]
print({nm: round(ms[i] / max(n[i], 1) * 1e3, 1) for i, nm in enumerate(names) if n[i]}, "us per launch")
