"""Weight-gradient products at the ViT-Tiny training shapes (K = 524 288 tokens): us per launch, TFLOP/s, fraction of the HBM roofline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi
lib = _capi.lib()
S = lambda: torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
K = 524288
for (M, N, cs) in ((192, 768, 1), (768, 192, 0), (192, 192, 0), (576, 192, 0), (192, 128, 1), (384, 384, 1), (1152, 384, 0)):
    A = torch.randn(K, M, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16()
    C = torch.zeros(M, N, device="cuda"); c = torch.zeros(M, device="cuda")
    f = lambda: lib.vitmarl_debug_gemm_dw(S(), M, N, K, A.data_ptr(), B.data_ptr(), C.data_ptr(), c.data_ptr() if cs else None, 0)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    by = K * (M + N) * 2
    print(f"dW [{M:4d} x {N:4d}] colsum={cs}: {t*1e6:7.1f} us  {2.0*M*N*K/t/1e12:6.0f} TF/s  {by/t/1e9:6.0f} GB/s = {by/t/1e9/peak:.2f} of HBM", flush=True)
    del A, B
