"""One ViT-S/16 training forward + backward at B = 8192, depth 1 (per-launch list for profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit
cfg = vit.ViTConfig(128, 128, 2, 16, 384, 1, 6, 1536)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
B = 8192
x = (torch.rand(B, 128, 128, 2, device="cuda") < 0.3).to(torch.bfloat16)
dy = torch.randn(B, cfg.dim, device="cuda")
for _ in range(3):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
torch.cuda.synchronize()
print("ok")
