"""Small driver for profiling: a few ViT forward passes (inference path) at B images."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = vit.ViTConfig(64, 64, 2, 8, 192, depth, 3, 768)
enc = vit.ViTEncoder(cfg)
params = vit.init_params(cfg, 0, "cuda")
packed = vit.pack_params(cfg, params)
x = (torch.rand(B, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
for _ in range(iters):
    y = enc.apply_packed(packed, x)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()))
