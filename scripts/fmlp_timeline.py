import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit, _capi
lib = _capi.lib()
cfg = vit.ViTConfig(64, 64, 2, 8, 192, 1, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
x = (torch.rand(4096, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
for _ in range(2): enc.apply_packed(packed, x)
lib.vitmarl_debug_fused_mlp_timeline(buf.data_ptr())
enc.apply_packed(packed, x); torch.cuda.synchronize()
lib.vitmarl_debug_fused_mlp_timeline(None)
t = buf.cpu().tolist(); t0 = t[0]
r = lambda i: t[i] - t0 if t[i] else None
print("compute: tile_start 0, XFULL", r(1), "LN done", r(2))
for c in range(12):
    print(f" c{c:2d}: acc1full {r(10+4*c)} gelu_done {r(11+4*c)} hempty {r(12+4*c)} hready {r(13+4*c)} | mma: fc1 start {r(100+4*c)} issued {r(101+4*c)} fc2: hready {r(102+4*c)} w2full {r(103+4*c)}")
print("acc2full", r(60), "tile end", r(61), "mma xnready", r(99))
