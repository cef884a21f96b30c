"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this pool: "find a bad access with bounds checks and asserts of
your own, small cases, and a comparison with the CPU reference").  Every output buffer of the kernels below is carved out of an arena
with sentinel-filled guard bands on both sides; after the launch the bands must be untouched.  Shapes are the awkward ones: ragged
row tiles (odd image counts), E not a multiple of the warps per CTA, M not a multiple of 32."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import synthetic_case                      # noqa: E402
from vitmarl_b200 import _capi, actor_critic, vit          # noqa: E402
from vitmarl_b200 import env as venv                       # noqa: E402
from vitmarl_b200.config import World_EnvironmentConfig    # noqa: E402

GUARD = 4096          # bytes on each side
PATTERN = 0xA5


class Arena:
    def __init__(self):
        self.bufs = []

    def alloc(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        pad = (-n) % 256
        raw = torch.full((GUARD + n + pad + GUARD,), PATTERN, dtype=torch.uint8, device="cuda")
        t = raw[GUARD:GUARD + n].view(dtype).view(shape)
        self.bufs.append((raw, n))
        return t

    def like(self, t):
        g = self.alloc(tuple(t.shape), t.dtype)
        g.copy_(t)
        return g

    def check(self, what):
        torch.cuda.synchronize()
        for i, (raw, n) in enumerate(self.bufs):
            lo, hi = raw[:GUARD], raw[GUARD + n + ((-n) % 256):]
            assert bool((lo == PATTERN).all()) and bool((hi == PATTERN).all()), f"{what}: guard band of buffer {i} overwritten"
            # (the 256-byte alignment pad directly behind the tensor is part of the band)
            assert bool((raw[GUARD + n:GUARD + n + ((-n) % 256)] == PATTERN).all()), f"{what}: bytes right behind buffer {i} overwritten"


def test_env_step_outputs_stay_in_bounds():
    cfg = World_EnvironmentConfig()
    for E, M in ((37, 13), (5, 113), (130, 9)):
        asks, bids, blocks = synthetic_case(E, M, steps=2, seed=E)
        A = Arena()
        st = venv.reset(cfg, torch.from_numpy(asks).cuda(), torch.from_numpy(bids).cuda(), M)
        st = venv.BookState(A.like(st.ask_raw_orders), A.like(st.bid_raw_orders), A.like(st.trades), A.like(st.best_asks), A.like(st.best_bids),
                            A.like(st.mid_price))
        bufs = venv.StepBuffers()
        kw = dict(n_levels=10, want_obs=True, want_raw=True, image_hw=(64, 64), image_patch=8, stat_agent_ids=[-100, -7], buffers=bufs)
        st, out = venv.step(cfg, st, torch.from_numpy(blocks[0]).cuda(), **kw)
        bufs.t = {k: A.like(v) for k, v in bufs.t.items()}          # every step output now sits between guard bands
        st, out = venv.step(cfg, st, torch.from_numpy(blocks[1]).cuda(), **kw)
        A.check(f"env_step E={E} M={M}")


@pytest.mark.parametrize("B,heads", [(1, 3), (3, 6), (149, 3)])
def test_attention_core_stays_in_bounds(B, heads):
    lib, S = _capi.lib(), torch.cuda.current_stream().cuda_stream
    D = heads * 64
    A = Arena()
    qkv = A.like(torch.randn(B * 64, 3 * D, device="cuda").bfloat16())
    dout = A.like(torch.randn(B * 64, D, device="cuda").bfloat16())
    out, dqkv = A.alloc((B * 64, D), torch.bfloat16), A.alloc((B * 64, 3 * D), torch.bfloat16)
    _capi.check(lib.vitmarl_attention_fwd(S, B, heads, qkv.data_ptr(), out.data_ptr()))
    _capi.check(lib.vitmarl_attention_bwd(S, B, heads, qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr()))
    A.check(f"attention B={B}")


@pytest.mark.parametrize("cfg,B", [(vit.VIT_PARITY, 1), (vit.VIT_PARITY, 37), (vit.ViTConfig(128, 128, 2, 16, 384, 1, 6, 1536), 3)])
def test_encoder_forward_backward_stay_in_bounds(cfg, B):
    """Workspace, output, gradient table and dx between guard bands: inference (fused blocks), training forward, backward
    (fused MLP backward at D = 192, unfused elsewhere)."""
    A = Arena()
    enc = vit.ViTEncoder(cfg)
    packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
    shape = enc._shape(B)
    need = max(_capi.lib().vitmarl_vit_workspace_bytes(ctypes.byref(shape), 1), _capi.lib().vitmarl_vit_workspace_bytes(ctypes.byref(shape), 0))
    enc._ws = A.alloc((need,), torch.uint8)
    x = (torch.rand(B, cfg.img_h, cfg.img_w, cfg.channels, device="cuda") < 0.3).to(torch.bfloat16)
    y = A.alloc((B, cfg.dim), torch.float32)
    enc.apply_packed(packed, x, out=y)
    A.check("vit inference")
    enc.apply_packed(packed, x, train=True, out=y)
    grads = [A.alloc(tuple(t.shape), torch.float32) for t in packed]
    _, dx = enc.vjp_packed(packed, torch.randn(B, cfg.dim, device="cuda"), grads=grads, want_dx=True)
    A.check("vit train forward + backward")
    assert all(bool(torch.isfinite(g).all()) for g in grads) and bool(torch.isfinite(dx.float()).all())


def test_policy_head_and_gae_stay_in_bounds():
    lib, S = _capi.lib(), torch.cuda.current_stream().cuda_stream
    A = Arena()
    R, K0, K1, N = 77, 23, 192, 130
    x0, x1 = torch.randn(R, K0, device="cuda"), torch.randn(R, K1, device="cuda")
    W, b = torch.randn(K0 + K1, N, device="cuda"), torch.randn(N, device="cuda")
    y = A.alloc((R, N), torch.float32)
    _capi.check(lib.vitmarl_dense_f32(S, R, K0, K1, N, x0.data_ptr(), K0, x1.data_ptr(), K1, W.data_ptr(), b.data_ptr(), 1, y.data_ptr(), N))
    H = 128
    gi, gh, bh, h = torch.randn(R, 3 * H, device="cuda"), torch.randn(R, 3 * H, device="cuda"), torch.randn(H, device="cuda"), torch.randn(R, H, device="cuda")
    ho = A.alloc((R, H), torch.float32)
    _capi.check(lib.vitmarl_gru_cell_f32(S, R, H, gi.data_ptr(), gh.data_ptr(), bh.data_ptr(), h.data_ptr(), None, ho.data_ptr()))
    Sx, B = 13, 301
    adv, tgt = A.alloc((Sx, B), torch.float32), A.alloc((Sx, B), torch.float32)
    r, v, d, lv = torch.randn(Sx, B, device="cuda"), torch.randn(Sx, B, device="cuda"), torch.zeros(Sx, B, dtype=torch.uint8, device="cuda"), torch.randn(B, device="cuda")
    _capi.check(lib.vitmarl_gae_f32(S, Sx, B, 0.99, 0.95, r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), adv.data_ptr(), tgt.data_ptr()))
    A.check("policy head / gae")
    assert torch.allclose(y, torch.relu(torch.cat([x0, x1], 1) @ W + b), atol=1e-3)
