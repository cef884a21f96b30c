"""clock64 timeline of the fused MLP backward kernel (CTA 0, its second tile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit, _capi
cfg = vit.ViTConfig(64, 64, 2, 8, 192, 1, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
B = 8192
x = (torch.rand(B, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
dy = torch.randn(B, cfg.dim, device="cuda")
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(2):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
enc.options.debug_timeline = buf.data_ptr()
enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy); torch.cuda.synchronize()
enc.options.debug_timeline = None
t = buf.cpu().tolist(); t0 = t[0]
r = lambda i: (t[i] - t0) if t[i] else None
print("tile: start 0, x landed", r(1), "LN done", r(2), "xhat stored", r(3))
for c in range(12):
    print(f" c{c:2d} epi: wait_acc {r(10+4*c)} got {r(11+4*c)} hready {r(12+4*c)} staged {r(13+4*c)} | issuer: fc1g start {r(100+6*c)} a1free {r(101+6*c)} w1full {r(102+6*c)} w2full {r(103+6*c)} | X: wait_h {r(104+6*c)} go {r(105+6*c)}")
for c in (3, 4, 8):
    b = t[11 + 4 * c]
    print(f" c{c} detail (from acc got): ld_done {t[200+8*c]-b} math_done {t[201+8*c]-b} st_done {t[202+8*c]-b} hready {t[12+4*c]-b} stgfree {t[203+8*c]-b} sts_done {t[204+8*c]-b} fence_done {t[205+8*c]-b} staged(+colsum) {t[13+4*c]-b}")
print("final: wait acc3", r(70), "got", r(71), "done", r(72))
tm = _capi.Timing(); enc.options.timing = tm.handle
for _ in range(3):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
ms, n, _ = tm.read()
print({nm: round(ms[i] / max(n[i], 1) * 1e3, 1) for i, nm in enumerate(_capi.Timing.NAMES) if n[i]}, "us per launch")
