"""NumPy restatement == C restatement on adversarial and synthetic streams (CPU only)."""
import numpy as np
import pytest

from oracle import lob_oracle as O
from tests.util import fuzz_case, synthetic_case


def _numpy_step(asks, bids, trades, msgs, n_keep):
    outs = [], [], [], [], []
    for e in range(asks.shape[0]):
        (a, b, t), (ba, bb) = O.scan_through_entire_array_save_bidask(msgs[e], (asks[e], bids[e], trades[e]), n_keep)
        for lst, v in zip(outs, (a, b, t, ba, bb)):
            lst.append(v)
    return [np.stack(x) for x in outs]


@pytest.mark.parametrize("seed,N,T,M", [(0, 5, 3, 40), (1, 12, 4, 60), (2, 33, 8, 50), (3, 64, 100, 30), (4, 100, 100, 20)])
def test_fuzz_step_numpy_vs_c(c_oracle, seed, N, T, M):
    rng = np.random.default_rng(seed)
    asks, bids, trades, msgs = fuzz_case(rng, 12, N, T, M)
    want = _numpy_step(asks, bids, trades, msgs, M)
    got = c_oracle.lob_step(asks, bids, msgs, trades_in=trades)
    for w, g, name in zip(want, got, ("asks", "bids", "trades", "best_asks", "best_bids")):
        assert np.array_equal(w, g), name
    # n_keep < M keeps the LAST n_keep rows (JOBA:752)
    got2 = c_oracle.lob_step(asks, bids, msgs, trades_in=trades, n_keep=7)
    assert np.array_equal(got2[3], want[3][:, -7:]) and np.array_equal(got2[4], want[4][:, -7:])


def test_synthetic_step_and_render_numpy_vs_c(c_oracle):
    asks, bids, blocks = synthetic_case(6, 40, steps=2)
    trades = np.full((6, 100, 8), -1, np.int32)
    for msgs in blocks:
        want = _numpy_step(asks, bids, trades, msgs, 40)
        got = c_oracle.lob_step(asks, bids, msgs)
        for w, g in zip(want, got):
            assert np.array_equal(w, g)
        asks, bids = got[0], got[1]
        assert (got[2][:, :, 0] != -1).any(), "synthetic stream should produce trades"
    fa, fb, mid = c_oracle.ffill_mid(got[3], got[4], got[3][:, 0, 0], got[4][:, 0, 0])
    raw, norm, img = c_oracle.render(asks, bids, mid_price=mid, n_levels=10, tick=100, H=64, W=64)
    for e in range(6):
        assert np.array_equal(fa[e], O.ffill_best_prices(got[3][e], got[3][e, 0, 0]))
        assert mid[e] == O.mid_price_f32(fb[e, -1, 0], fa[e, -1, 0])
        r = O.get_vision_L2_state(asks[e], bids[e], 10)
        assert np.array_equal(raw[e], r)
        assert norm[e].tobytes() == O.normalize_vision_obs(r, mid[e], 100).tobytes()
        assert np.array_equal(img[e], O.render_image(asks[e], bids[e], 64, 64, 100))


@pytest.mark.parametrize("seed", [10, 11])
def test_fuzz_render_numpy_vs_c(c_oracle, seed):
    rng = np.random.default_rng(seed)
    asks, bids, _, _ = fuzz_case(rng, 10, 20, 4, 1)
    asks[..., 0] = np.where(asks[..., 0] != -1, asks[..., 0] * 100, -1)
    bids[..., 0] = np.where(bids[..., 0] != -1, bids[..., 0] * 100, -1)
    mid = rng.integers(9900, 10900, size=10).astype(np.float32) + np.float32(0.5)
    raw, norm, img = c_oracle.render(asks, bids, mid_price=mid, n_levels=7, tick=100, H=16, W=32)
    for e in range(10):
        r = O.get_vision_L2_state(asks[e], bids[e], 7)
        assert np.array_equal(raw[e], r)
        assert norm[e].tobytes() == O.normalize_vision_obs(r, mid[e], 100).tobytes()
        assert np.array_equal(img[e], O.render_image(asks[e], bids[e], 16, 32, 100))
