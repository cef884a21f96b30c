"""Bring-up check of the tcgen05 GEMM against torch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitmarl_b200 import _capi
lib = _capi.lib()
S = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)

def run(M, N, K, epi=0, bias=False, res=False, pos=0, a_mn=False, b_mn=False):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    bias_t = torch.randn(N, device="cuda") if bias else None
    res_t = torch.randn(M, N, device="cuda").bfloat16() if res else None
    pos_t = torch.randn(pos, N, device="cuda") if pos else None
    ref = A.float() @ B.float().t()
    if bias: ref = ref + bias_t
    if epi == 1: ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if pos: ref = ref + pos_t.repeat((M + pos - 1) // pos, 1)[:M]
    if res: ref = ref + res_t.float()
    if epi in (2, 3):
        C = torch.zeros(M, N, device="cuda", dtype=torch.float32) + (1.0 if epi == 3 else 0.0)
        if epi == 3: ref = 1.0 + 0.5 * ref
    else:
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rc = lib.vitmarl_gemm_bf16(S(), M, N, K, Am.data_ptr(), Am.stride(0), int(a_mn), Bm.data_ptr(), Bm.stride(0), int(b_mn),
                               C.data_ptr(), N, epi, bias_t.data_ptr() if bias else None, res_t.data_ptr() if res else None, N,
                               pos_t.data_ptr() if pos else None, pos, 0.5 if epi == 3 else 1.0)
    torch.cuda.synchronize()
    assert rc == 0, (rc, lib.vitmarl_last_error())
    err = (C.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
    print(f"M={M} N={N} K={K} epi={epi} bias={bias} res={res} pos={pos} a_mn={a_mn} b_mn={b_mn}: rel-max-err {err:.2e}", flush=True)
    return err

errs = []
errs.append(run(128, 64, 64))
errs.append(run(256, 192, 192))
errs.append(run(1000, 576, 192, bias=True))
errs.append(run(4096, 768, 192, epi=1, bias=True))
errs.append(run(4096, 192, 768, bias=True, res=True))
errs.append(run(640, 128, 128, bias=True, pos=64))
errs.append(run(512, 320, 256, epi=2, bias=True))
errs.append(run(4096, 192, 576, b_mn=True))                 # dX = dY . W
errs.append(run(192, 192, 8192, epi=3, a_mn=True, b_mn=True))   # dW = dY^T X (split-K, atomics)
errs.append(run(768, 192, 16384, epi=3, a_mn=True, b_mn=True))
errs.append(run(256, 128, 512, a_mn=True))
# CTA-pair kernel with MN-major operands
errs.append(run(4096, 384, 1536, b_mn=True, res=True))          # staged epilogue, 1.5-block B halves
errs.append(run(5000, 384, 3072, b_mn=True, bias=True))         # per-thread epilogue, ragged M
errs.append(run(1536, 384, 32768, epi=3, a_mn=True, b_mn=True)) # dW, 6 x 2 pair tiles, split-K
errs.append(run(1152, 384, 16384, epi=3, a_mn=True, b_mn=True)) # dW, last pair tile half empty
errs.append(run(2048, 192, 200 * 64, epi=3, a_mn=True, b_mn=True))
errs.append(run(384, 1536, 20000, epi=3, a_mn=True, b_mn=True))  # 256 x 384 work items, operands swapped + transposed output
errs.append(run(1280, 768, 9984, epi=3, a_mn=True, b_mn=True))   # two 384-column work items per row block
# bias gradient (column sums of A) fused into the weight-gradient kernel / separate pass for the other routes
def run_colsum(M, N, K):
    A = (torch.randn(K, M, device="cuda") * 0.5).bfloat16(); B = (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
    C = torch.zeros(M, N, device="cuda"); cs = torch.full((M,), 2.0, device="cuda")
    rc = lib.vitmarl_debug_gemm_dw(S(), M, N, K, A.data_ptr(), B.data_ptr(), C.data_ptr(), cs.data_ptr())
    torch.cuda.synchronize()
    assert rc == 0, (rc, lib.vitmarl_last_error())
    e1 = (C - A.float().t() @ B.float()).abs().max().item() / (A.float().t() @ B.float()).abs().max().item()
    ref = 2.0 + A.float().sum(0)
    e2 = (cs - ref).abs().max().item() / ref.abs().max().item()
    print(f"dW+colsum M={M} N={N} K={K}: rel-max-err {e1:.2e} / colsum {e2:.2e}", flush=True)
    return max(e1, e2)
for shp in ((1536, 384, 32768), (1152, 384, 20000), (384, 384, 8192), (384, 1536, 8192), (576, 192, 4096), (768, 768, 6000)):
    errs.append(run_colsum(*shp))
assert max(errs) < 1e-2, max(errs)
# throughput
for (M, N, K) in ((262144, 576, 192), (262144, 768, 192), (262144, 192, 768)):
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for i in range(3): lib.vitmarl_gemm_bf16(S(), M, N, K, A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, 0, None, None, N, None, 0, 1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): lib.vitmarl_gemm_bf16(S(), M, N, K, A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, 0, None, None, N, None, 0, 1.0)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    print(f"GEMM {M}x{N}x{K}: {t*1e6:.0f} us  {2*M*N*K/t/1e12:.0f} TFLOP/s  {(M*K+M*N)*2/t/1e9:.0f} GB/s")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(10): torch.matmul(A, B.t(), out=C)
    t1.record(); torch.cuda.synchronize()
    print(f"   cuBLAS: {t0.elapsed_time(t1)/10*1e3:.0f} us")
print("GEMM OK")
