import sys; sys.path.insert(0, ".")
import torch
from oracle import vit_oracle as VO
from vitmarl_b200 import vit
from tests.test_vit_gpu import _images, _perturbed_params
for cfg, B in ((vit.VIT_SMALL_16, 296), (vit.VIT_TINY_8, 296)):
    for fused in (1, 0):
        if fused == 0 and cfg.dim != 192: continue
        params = _perturbed_params(cfg, 3); x = _images(B, cfg, 1)
        dy = torch.randn(B, cfg.dim, generator=torch.Generator().manual_seed(1)).cuda()
        enc = vit.ViTEncoder(cfg); enc.options.fused = fused
        enc.apply({"params": params}, x, train=True)
        grads = enc.vjp({"params": params}, dy)
        _, ref, _ = VO.vit_value_and_grad(cfg, params, x, dy)
        worst_cos, worst_l2 = 1.0, 0.0
        for (name, g), (_, r) in zip(VO.tree_leaves(grads), VO.tree_leaves(ref)):
            if name.endswith("key/bias"): continue
            g, r = g.float().reshape(-1), r.float().reshape(-1)
            cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30)); l2 = float((g - r).norm() / (r.norm() + 1e-30))
            if l2 > worst_l2: worst_l2, wn = l2, name
            worst_cos = min(worst_cos, cos)
        print(f"dim {cfg.dim} fused={fused}: worst cos {worst_cos:.6f} worst rel-L2 {worst_l2:.4f} ({wn})")
