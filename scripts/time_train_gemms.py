"""Training-path GEMM shapes of ViT-S/16 at B=8192 (tokens M = 524288, D = 384): 1-CTA vs CTA-pair kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi
lib = _capi.lib()
S = lambda: torch.cuda.current_stream().cuda_stream
T, D = 524288, 384
FLAGS = 0

def timeit(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

def dx(n_out, n_in):      # dX[T, n_in] = dY[T, n_out] . W[n_out, n_in]   (B = W read MN-major: [K = n_out][N = n_in])
    dY = torch.randn(T, n_out, device="cuda").bfloat16(); W = torch.randn(n_out, n_in, device="cuda").bfloat16()
    C = torch.empty(T, n_in, device="cuda", dtype=torch.bfloat16)
    f = lambda: lib.vitmarl_gemm_bf16(S(), T, n_in, n_out, dY.data_ptr(), n_out, 0, W.data_ptr(), n_in, 1, C.data_ptr(), n_in, 0, None, None, n_in, None, 0, 1.0, FLAGS)
    return f, 2.0 * T * n_in * n_out, (T * n_out + T * n_in) * 2

def dw(n_out, n_in):      # dW[n_out, n_in] += dY^T[n_out, T] . X[T, n_in]    (both MN-major, fp32 red.add)
    dY = torch.randn(T, n_out, device="cuda").bfloat16(); X = torch.randn(T, n_in, device="cuda").bfloat16()
    C = torch.zeros(n_out, n_in, device="cuda", dtype=torch.float32)
    f = lambda: lib.vitmarl_gemm_bf16(S(), n_out, n_in, T, dY.data_ptr(), n_out, 1, X.data_ptr(), n_in, 1, C.data_ptr(), n_in, 3, None, None, n_in, None, 0, 1.0, FLAGS)
    return f, 2.0 * T * n_in * n_out, (T * n_out + T * n_in) * 2

for name, mk, shapes in (("dX", dx, ((1152, 384), (384, 384), (1536, 384), (384, 1536))), ("dW", dw, ((1152, 384), (384, 384), (1536, 384), (384, 1536)))):
    for (no, ni) in shapes:
        f, fl, by = mk(no, ni)
        out = []
        for two in (0, 1):
            FLAGS = 0 if two else _capi.GEMM_NO_2CTA
            t = timeit(f)
            out.append(f"{'pair' if two else '1cta'} {t*1e6:7.0f} us {fl/t/1e12:6.0f} TF/s {by/t/1e9:6.0f} GB/s")
        print(f"{name} out={no:5d} in={ni:5d}: " + " | ".join(out), flush=True)
