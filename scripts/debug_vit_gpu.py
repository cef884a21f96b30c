"""Per-tensor error report of the ViT forward/backward vs the fp32 oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as VO
from vitmarl_b200 import vit
from tests.test_vit_gpu import _images, _perturbed_params, _rel

for cfg, B in [(vit.ViTConfig(64, 64, 2, 8, 192, 0, 3, 768), 6), (vit.ViTConfig(64, 64, 2, 8, 192, 1, 3, 768), 8),
               (vit.VIT_PARITY, 16), (vit.ViTConfig(128, 128, 2, 16, 384, 2, 6, 1536), 4), (vit.VIT_TINY_8, 32)]:
    params = _perturbed_params(cfg, 3)
    x = _images(B, cfg, 1)
    dy = torch.randn(B, cfg.dim, generator=torch.Generator().manual_seed(1)).cuda()
    enc = vit.ViTEncoder(cfg)
    y = enc.apply({"params": params}, x, train=True)
    grads, dx = enc.vjp({"params": params}, dy, want_dx=True)
    torch.cuda.synchronize()
    yr, ref, ref_dx = VO.vit_value_and_grad(cfg, params, x, dy, want_dx=True)
    print(f"== depth {cfg.depth} dim {cfg.dim} B {B}: fwd rel err {_rel(y, yr):.3e}")
    for (name, g), (_, r) in zip(VO.tree_leaves(grads), VO.tree_leaves(ref)):
        g, r = g.float().reshape(-1), r.float().reshape(-1)
        cos = torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)
        l2 = (g - r).norm() / (r.norm() + 1e-30)
        flag = "" if (cos >= 0.999 and l2 <= 3e-2) else "   <-- BAD"
        print(f"   {name:70s} cos {cos:.5f} relL2 {l2:.3e}{flag}")
    g, r = dx.float().reshape(-1), ref_dx.reshape(-1)
    print(f"   dx cos {torch.dot(g, r) / (g.norm() * r.norm()):.5f}")
