import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi
lib = _capi.lib(); S = torch.cuda.current_stream().cuda_stream
for (M, N, K) in ((8192, 3840, 4096), (8192, 4096, 4096), (8192, 4160, 4096)):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    f = lambda: lib.vitmarl_gemm_bf16(S, M, N, K, A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, 0, None, None, N, None, 0, 1.0, 0)
    for i in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    bn = 192 if N % 192 == 0 else (128 if N % 128 == 0 else 64)
    per_mma_cycles = t * 1.9e9 / ((M / 128) * (N / bn) * (K / 16) / 148)
    print(f"GEMM {M}x{N}x{K} BN={bn}: {t*1e6:.0f} us  {2*M*N*K/t/1e12:.0f} TFLOP/s   ~{per_mma_cycles:.0f} cycles per MMA (floor {128*bn//256})")
