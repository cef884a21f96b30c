"""Per-kernel-class time of one training forward + backward (VitmarlVitOptions.timing), ViT-S/16 or ViT-Tiny/8 at depth L."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit, _capi
model = sys.argv[1] if len(sys.argv) > 1 else "small16"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
cfg = vit.ViTConfig(128, 128, 2, 16, 384, L, 6, 1536) if model == "small16" else vit.ViTConfig(64, 64, 2, 8, 192, L, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
x = (torch.rand(B, cfg.img_h, cfg.img_w, 2, device="cuda") < 0.3).to(torch.bfloat16)
dy = torch.randn(B, cfg.dim, device="cuda")
for _ in range(2):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
tm = _capi.Timing(); enc.options.timing = tm.handle
reps = 3
for _ in range(reps):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
torch.cuda.synchronize()
ms, n, _ = tm.read()
print(model, "L", L, "B", B, {nm: (round(ms[i] / reps, 3), n[i] // reps) for i, nm in enumerate(_capi.Timing.NAMES) if n[i]}, "(ms per fwd+bwd, launches)")
