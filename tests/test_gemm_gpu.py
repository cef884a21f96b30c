"""tcgen05 GEMM family against a torch fp32 reference (floating point: tolerance 1e-2 of the output's max magnitude for bf16
outputs, 1e-4 for the fp32 split-K accumulations): the 1-CTA kernel, the CTA-pair kernel with K-major and MN-major operands,
and the CTA-pair weight-gradient kernel (256 x 384 work items, operand swap + transposed output, fused bias gradient)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from vitmarl_b200 import _capi  # noqa: E402


def _S():
    return torch.cuda.current_stream().cuda_stream


def _gemm(M, N, K, *, epi=0, bias=False, res=False, a_mn=False, b_mn=False, seed=0, flags=0):
    lib = _capi.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    bias_t = torch.randn(N, device="cuda", generator=g) if bias else None
    res_t = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bias_t
    if epi == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if res:
        ref = ref + res_t.float()
    if epi == 3:                                   # fp32 accumulate-into (split-K reductions), scaled
        C = torch.ones(M, N, device="cuda", dtype=torch.float32)
        ref = 1.0 + 0.5 * ref
    else:
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rc = lib.vitmarl_gemm_bf16(_S(), M, N, K, Am.data_ptr(), Am.stride(0), int(a_mn), Bm.data_ptr(), Bm.stride(0), int(b_mn),
                               C.data_ptr(), N, epi, bias_t.data_ptr() if bias else None, res_t.data_ptr() if res else None, N,
                               None, 0, 0.5 if epi == 3 else 1.0, flags)
    torch.cuda.synchronize()
    _capi.check(rc)
    return (C.float() - ref).abs().max().item() / ref.abs().max().item()


@pytest.mark.parametrize("two_cta", [1, 0])
def test_forward_and_dx_shapes(two_cta):
    f = 0 if two_cta else _capi.GEMM_NO_2CTA       # per-call flag (the library has no global switch)
    assert _gemm(1000, 576, 192, bias=True, flags=f) < 1e-2                       # ragged M
    assert _gemm(2048, 768, 192, epi=1, bias=True, flags=f) < 1e-2                # bias + GELU
    assert _gemm(2048, 192, 768, bias=True, res=True, flags=f) < 1e-2             # residual epilogue
    assert _gemm(2048, 384, 1536, b_mn=True, res=True, flags=f) < 1e-2            # dX: MN-major B, 1.5-block halves on the pair kernel
    assert _gemm(1300, 384, 3072, b_mn=True, bias=True, flags=f) < 1e-2           # large K: per-thread epilogue, ragged M


@pytest.mark.parametrize("M,N,K", [(192, 192, 4096), (768, 192, 8192), (1536, 384, 16384), (1152, 384, 10000),
                                   (384, 1536, 8192), (1280, 768, 6016), (384, 384, 4096)])
def test_weight_gradient_products(M, N, K):
    """dW = dY^T . X over the token dimension, both operands MN-major, split-K with fp32 reductions."""
    assert _gemm(M, N, K, epi=3, a_mn=True, b_mn=True) < 1e-4


@pytest.mark.parametrize("M,N,K", [(1536, 384, 8192), (1152, 384, 5000), (384, 384, 4096), (384, 1536, 4096), (576, 192, 2048),
                                   (768, 768, 3000)])
def test_weight_gradient_with_fused_bias_gradient(M, N, K):
    lib = _capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = (torch.randn(K, M, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(K, N, device="cuda", generator=g) * 0.5).bfloat16()
    C = torch.zeros(M, N, device="cuda")
    cs = torch.full((M,), 2.0, device="cuda")
    _capi.check(lib.vitmarl_debug_gemm_dw(_S(), M, N, K, A.data_ptr(), B.data_ptr(), C.data_ptr(), cs.data_ptr(), 0))
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()
    assert (C - ref).abs().max().item() / ref.abs().max().item() < 1e-4
    refs = 2.0 + A.float().sum(0)
    assert (cs - refs).abs().max().item() / refs.abs().max().item() < 1e-4
