"""One ViT-Tiny/8 training forward + backward at B = 8192 (fused MLP path) -- the command line profiled for profiles/r02_*train*."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit
cfg = vit.ViTConfig(64, 64, 2, 8, 192, 2, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
B = 8192
x = (torch.rand(B, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
dy = torch.randn(B, cfg.dim, device="cuda")
for _ in range(3):
    enc.apply_packed(packed, x, train=True); enc.vjp_packed(packed, dy)
torch.cuda.synchronize()
print("ok")
