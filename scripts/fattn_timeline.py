import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import vit, _capi
lib = _capi.lib()
cfg = vit.ViTConfig(64, 64, 2, 8, 192, 1, 3, 768)
enc = vit.ViTEncoder(cfg)
packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
x = (torch.rand(4096, 64, 64, 2, device="cuda") < 0.3).to(torch.bfloat16)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(2): enc.apply_packed(packed, x)
lib.vitmarl_debug_fused_mlp_timeline(buf.data_ptr())
enc.apply_packed(packed, x); torch.cuda.synchronize()
lib.vitmarl_debug_fused_mlp_timeline(None)
t = buf.cpu().tolist()[256:]; t0 = t[0]
r = lambda i: t[i] - t0 if t[i] else None
print("compute: XFULL", r(1), "LN done", r(2), "| mma xnready", r(99), "QKV0 issued", r(100))
for h in range(3):
    b = 8 * h
    print(f" h{h}: C: qkvfull {r(10+b)} epi_done {r(11+b)} sfull {r(12+b)} p_done {r(13+b)} ofull {r(14+b)} o_done {r(15+b)} | M: qkready {r(101+b)} S_issued {r(102+b)} nextQKV_issued {r(103+b)} pready {r(104+b)} PV_issued {r(105+b)}")
print("proj start", r(130), "issued", r(131), "| C: pfull", r(60), "tile end", r(61))
