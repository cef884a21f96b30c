"""Attention core (csrc/attention_tc.cu) alone at the training shapes: us per launch and fraction of the HBM roofline
(forward: read qkv, write out = 4 x [M, D] bf16; backward: read qkv + dout, write dqkv = 7 x [M, D])."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi
lib = _capi.lib()
S = lambda: torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for B, heads in ((8192, 3), (8192, 6)):
    D = heads * 64
    qkv = torch.randn(B * 64, 3 * D, device="cuda").bfloat16()
    dout = torch.randn(B * 64, D, device="cuda").bfloat16()
    out = torch.empty(B * 64, D, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    for name, fn, units in (("fwd", lambda: lib.vitmarl_attention_fwd(S(), B, heads, qkv.data_ptr(), out.data_ptr()), 4),
                            ("bwd", lambda: lib.vitmarl_attention_bwd(S(), B, heads, qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr()), 7)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        by = units * B * 64 * D * 2
        print(f"attention {name} B={B} heads={heads}: {t*1e6:8.1f} us  {by/t/1e9:7.0f} GB/s algorithmic = {by/t/1e9/peak:.2f} of measured HBM copy", flush=True)
