"""tcgen05 attention core (csrc/attention_tc.cu) against a torch fp32 reference of the same op: softmax(q k^T / 8) v per
(image, head) over 64 tokens, head dim 64, and its VJP.  bf16 operands / fp32 accumulation: forward within 1e-2 of the output's
max magnitude, gradients cosine >= 0.9995 and rel-L2 <= 2e-2 per tensor (P and dz are rounded to bf16 before the second
products, as on every tensor-core attention backward)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from vitmarl_b200 import _capi  # noqa: E402


def _ref(qkv, dout, B, heads):
    D = heads * 64
    x = qkv.float().reshape(B, 64, 3, heads, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)   # [3, B, h, T, d]
    q, k, v = x[0], x[1], x[2]
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * 64, D)
    (o * dout.float()).sum().backward()
    g = x.grad.permute(1, 3, 0, 2, 4).reshape(B * 64, 3 * D)
    return o.detach(), g


@pytest.mark.parametrize("B,heads", [(2, 3), (1, 3), (7, 6), (300, 3), (149, 6)])
def test_attention_core_forward_and_backward(B, heads):
    lib = _capi.lib()
    S = torch.cuda.current_stream().cuda_stream
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(B * 10 + heads)
    qkv = (torch.randn(B * 64, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * 64, D, device="cuda", generator=g).bfloat16()
    out = torch.full((B * 64, D), 7.0, device="cuda").bfloat16()
    dqkv = torch.full((B * 64, 3 * D), 7.0, device="cuda").bfloat16()
    _capi.check(lib.vitmarl_attention_fwd(S, B, heads, qkv.data_ptr(), out.data_ptr()))
    _capi.check(lib.vitmarl_attention_bwd(S, B, heads, qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr()))
    torch.cuda.synchronize()
    o_ref, g_ref = _ref(qkv, dout, B, heads)
    assert (out.float() - o_ref).abs().max().item() <= 1e-2 * o_ref.abs().max().item()
    for name, lo in (("dq", 0), ("dk", D), ("dv", 2 * D)):
        a, b = dqkv[:, lo:lo + D].float().reshape(-1), g_ref[:, lo:lo + D].reshape(-1)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        l2 = float((a - b).norm() / b.norm())
        assert cos >= 0.9995 and l2 <= 2e-2, (name, cos, l2)
    # bitwise deterministic (no atomics anywhere in these kernels)
    out2 = torch.empty_like(out)
    dqkv2 = torch.empty_like(dqkv)
    _capi.check(lib.vitmarl_attention_fwd(S, B, heads, qkv.data_ptr(), out2.data_ptr()))
    _capi.check(lib.vitmarl_attention_bwd(S, B, heads, qkv.data_ptr(), dout.data_ptr(), dqkv2.data_ptr()))
    assert torch.equal(out, out2) and torch.equal(dqkv, dqkv2)
