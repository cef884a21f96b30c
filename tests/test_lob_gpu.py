"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.
Run with `-m gpu` on a B200."""
import numpy as np
import pytest
import torch

from tests.util import fuzz_case, lobster_to_msg, synthetic_case

pytestmark = pytest.mark.gpu

from vitmarl_b200 import env as venv            # noqa: E402
from vitmarl_b200 import _capi, jaxob, vision    # noqa: E402
from vitmarl_b200.config import World_EnvironmentConfig  # noqa: E402

CFG = World_EnvironmentConfig()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def _cfg(T):
    import dataclasses
    return dataclasses.replace(CFG, nTrades=T, nTradesLogged=T)


def _check_step(C, asks, bids, trades, msgs, n_keep, T):
    want = C.lob_step(asks, bids, msgs, trades_in=trades, T=T, n_keep=n_keep)
    (a, b, t), (ba, bb) = jaxob.scan_through_entire_array_save_bidask(
        _cfg(T), None, dev(msgs), (dev(asks), dev(bids), None if trades is None else dev(trades)), n_keep)
    torch.cuda.synchronize()
    for w, g, name in zip(want, (a, b, t, ba, bb), ("asks", "bids", "trades", "best_asks", "best_bids")):
        g = host(g)
        assert g.shape == w.shape, name
        if not np.array_equal(w, g):
            bad = np.argwhere((w != g).reshape(w.shape[0], -1).any(axis=1)).ravel()
            raise AssertionError(f"{name}: {len(bad)} envs differ, first env {bad[0]}")
    return want


def test_c1_parity_config(c_oracle):
    """BASELINE configs[0]: 16 envs, single execution agent (M = 1 + 8), N = T = 100."""
    asks, bids, blocks = synthetic_case(16, 9, steps=8)
    for msgs in blocks:
        want = _check_step(c_oracle, asks, bids, None, msgs, 9, 100)
        asks, bids = want[0], want[1]


def test_g1_through_cuda(golden, c_oracle):
    from oracle import lob_oracle as O
    rows = np.array(golden["book_rows"], dtype=np.int32)
    init = O.init_msgs_from_l2(rows[0], (34200, 17459617))[None]
    a = jaxob.init_orderside(100, 1)
    b = jaxob.init_orderside(100, 1)
    a, b, t = jaxob.scan_through_entire_array(CFG, None, dev(init), (a, b, None))
    assert np.array_equal(host(jaxob.get_L2_state(a, b, 10, CFG))[0], rows[0])
    for i in range(1, 5):
        m = dev(lobster_to_msg(golden["messages_lobster"][i])[None, None])
        a, b, t = jaxob.scan_through_entire_array(CFG, None, m, (a, b, t))
        assert np.array_equal(host(jaxob.get_L2_state(a, b, 10, CFG))[0], rows[i])
    # batched init recipe on the device == oracle recipe
    assert np.array_equal(host(jaxob.init_msgs_from_l2(CFG, dev(rows[:1]), (34200, 17459617))), init)


@pytest.mark.parametrize("seed,N,T,M", [(0, 5, 3, 40), (1, 12, 4, 70), (2, 33, 8, 50), (3, 64, 100, 33),
                                        (4, 100, 100, 32), (5, 101, 7, 45), (6, 200, 16, 64), (7, 256, 5, 20), (8, 1, 1, 9)])
def test_fuzz_step(c_oracle, seed, N, T, M):
    """Adversarial streams: every quirk (Q1-Q11), odd N (no TMA bulk path), N up to 256."""
    rng = np.random.default_rng(seed)
    asks, bids, trades, msgs = fuzz_case(rng, 97, N, T, M)
    _check_step(c_oracle, asks, bids, trades, msgs, M, T)
    _check_step(c_oracle, asks, bids, None, msgs, 5, T)
    _check_step(c_oracle, asks, bids, trades, msgs[:, :0], 0, T)      # M = 0: pure copy


@pytest.mark.parametrize("seed,N,T,M", [(20, 5, 3, 40), (21, 12, 4, 70), (22, 33, 8, 50), (23, 64, 100, 33),
                                        (24, 100, 100, 64), (25, 101, 7, 45), (26, 200, 16, 64), (27, 256, 5, 20),
                                        (28, 1, 1, 9), (29, 100, 6, 100)])
def test_fuzz_step_tidy_books(c_oracle, seed, N, T, M):
    """Same adversarial value ranges on WELL-FORMED books: the register fast path of the kernel (price-time priority
    with colliding prices / times / ids, Q2 full sides, Q4/Q5 cancels, trade-log overflow, prefix trade logs) and
    the hand-over to the literal path in the middle of a step."""
    rng = np.random.default_rng(seed)
    for fill in (0.3, 0.9):
        asks, bids, trades, msgs = fuzz_case(rng, 97, N, T, M, fill=fill, tidy=True)
        _check_step(c_oracle, asks, bids, trades, msgs, M, T)
        _check_step(c_oracle, asks, bids, None, msgs, min(5, M), T)
        # limit orders and non-negative cancels only: no environment ever leaves the fast path
        calm = msgs.copy()
        calm[..., 2] = np.abs(calm[..., 2])
        _check_step(c_oracle, asks, bids, None, calm, M, T)


def test_book_depth_sweep_and_large_m(c_oracle):
    """BASELINE configs[2] shapes at oracle-sized E: capacity 10..100, M = 100 and 113."""
    for N in (10, 20, 50, 100):
        asks, bids, blocks = synthetic_case(64, 100, steps=2, N=N)
        for msgs in blocks:
            want = _check_step(c_oracle, asks, bids, None, msgs, 100, 100)
            asks, bids = want[0], want[1]
    asks, bids, blocks = synthetic_case(32, 113, steps=1)
    _check_step(c_oracle, asks, bids, None, blocks[0], 113, 100)


def test_inplace_and_scan_without_bidask(c_oracle):
    asks, bids, blocks = synthetic_case(33, 13, steps=1)
    want = c_oracle.lob_step(asks, bids, blocks[0])
    a, b = dev(asks), dev(bids)
    (a2, b2, t2), _ = jaxob.scan_through_entire_array_save_bidask(CFG, None, dev(blocks[0]), (a, b, None), 13, inplace=True)
    assert a2.data_ptr() == a.data_ptr()
    assert np.array_equal(host(a2), want[0]) and np.array_equal(host(b2), want[1]) and np.array_equal(host(t2), want[2])
    a3, b3, t3 = jaxob.scan_through_entire_array(CFG, None, dev(blocks[0]), (dev(asks), dev(bids), None))
    assert np.array_equal(host(a3), want[0]) and np.array_equal(host(t3), want[2])


def test_best_bid_ask_and_render_fuzz(c_oracle):
    rng = np.random.default_rng(42)
    for N, n_levels, H, W in ((20, 7, 16, 32), (100, 10, 64, 64), (37, 32, 8, 8), (256, 3, 128, 128)):
        asks, bids, _, _ = fuzz_case(rng, 50, N, 4, 1)
        asks[..., 0] = np.where(asks[..., 0] != -1, asks[..., 0] * 100 + rng.integers(0, 2, asks.shape[:2]) * 50, -1)
        bids[..., 0] = np.where(bids[..., 0] != -1, bids[..., 0] * 100, -1)
        mid = (rng.integers(9900, 10900, size=50) + 0.5).astype(np.float32)
        raw_w, norm_w, img_w = c_oracle.render(asks, bids, mid_price=mid, n_levels=n_levels, tick=100, H=H, W=W)
        ba_w, bb_w = c_oracle.best_bid_ask(asks, bids)
        a, b = dev(asks), dev(bids)
        ba, bb = jaxob.get_best_bid_and_ask_inclQuants(CFG, a, b)
        assert np.array_equal(host(ba), ba_w) and np.array_equal(host(bb), bb_w)
        assert np.array_equal(host(jaxob.get_vision_L2_state(a, b, n_levels, CFG)), raw_w)
        l2 = host(jaxob.get_L2_state(a, b, n_levels, CFG)).reshape(50, n_levels, 4)
        assert np.array_equal(l2[:, :, 0], raw_w[:, :, 0, 0]) and np.array_equal(l2[:, :, 3], raw_w[:, :, 1, 1])
        norm = host(vision.normalize_vision_obs(a, b, dev(mid), n_levels, 100))
        assert norm.tobytes() == norm_w.tobytes()                      # bit-exact incl. log1p
        img8 = host(vision.render_image(a, b, H, W, 100, torch.uint8))
        assert np.array_equal(img8, img_w)
        img16 = vision.render_image(a, b, H, W, 100, torch.bfloat16)
        assert np.array_equal(host(img16.float()).astype(np.uint8), img_w)


def test_log1p_bit_exact_over_volume_range(c_oracle):
    """norm channel 1 = log1p(volume): sweep volumes through a one-level book."""
    vols = np.concatenate([np.arange(1, 4097), np.random.default_rng(0).integers(1, 2 ** 31 - 1, 4096)]).astype(np.int32)
    E = vols.shape[0]
    asks = np.full((E, 2, 6), -1, np.int32)
    asks[:, 0] = np.stack([np.full(E, 1000), vols, np.ones(E), np.ones(E), np.ones(E), np.ones(E)], axis=1)
    bids = np.full((E, 2, 6), -1, np.int32)
    mid = np.full((E,), 900, np.float32)
    _, norm_w, _ = c_oracle.render(asks, bids, mid_price=mid, n_levels=2, tick=100)
    norm = host(vision.normalize_vision_obs(dev(asks), dev(bids), dev(mid), 2, 100))
    assert norm.tobytes() == norm_w.tobytes()


def _oracle_env_step(C, asks, bids, msgs, last_a, last_b, n_levels, H, W):
    a, b, t, ba, bb = C.lob_step(asks, bids, msgs)
    fa, fb, mid = C.ffill_mid(ba, bb, last_a, last_b)
    raw, norm, img = C.render(a, b, mid_price=mid, n_levels=n_levels, tick=100, H=H, W=W)
    return a, b, t, fa, fb, mid, raw, norm, img


# (4096, 13, 8) is BASELINE configs[1] at its full size: 8 chained steps, every output leaf compared byte for byte with the C oracle
@pytest.mark.parametrize("E,M,steps", [(16, 9, 6), (257, 13, 4), (64, 113, 2), (4096, 13, 8)])
def test_fused_env_step_equals_composition(c_oracle, E, M, steps):
    asks, bids, blocks = synthetic_case(E, M, steps=steps)
    state = venv.reset(CFG, dev(asks), dev(bids), M)
    ba0, bb0 = c_oracle.best_bid_ask(asks, bids)
    assert np.array_equal(host(state.best_asks[:, -1]), ba0)
    assert host(state.mid_price).tobytes() == ((bb0[:, 0] + ba0[:, 0]).astype(np.float32) / np.float32(2)).tobytes()
    last_a, last_b = ba0[:, 0].copy(), bb0[:, 0].copy()
    for msgs in blocks:
        w = _oracle_env_step(c_oracle, asks, bids, msgs, last_a, last_b, 10, 64, 64)
        state, out = venv.step(CFG, state, dev(msgs), n_levels=10, want_obs=True, want_raw=True, image_hw=(64, 64),
                               image_dtype=torch.uint8)
        torch.cuda.synchronize()
        got = (state.ask_raw_orders, state.bid_raw_orders, state.trades, state.best_asks, state.best_bids,
               state.mid_price, out.vision_raw, out.vision_obs, out.image)
        for name, x, y in zip("asks bids trades best_asks best_bids mid raw norm image".split(), w, got):
            assert host(y).tobytes() == np.ascontiguousarray(x).tobytes(), name
        asks, bids, last_a, last_b = w[0], w[1], w[3][:, -1, 0].copy(), w[4][:, -1, 0].copy()


def test_raster_written_as_patch_matrix(c_oracle):
    """`image_patch=p`: the fused step writes the same raster in the ViT's patch-matrix order (token = py*(W/p)+px,
    feature = (ph, pw, c)) -- bit-identical to patchifying the [E,H,W,2] image -- and the encoder takes it as is."""
    from vitmarl_b200 import vit
    asks, bids, blocks = synthetic_case(70, 13, steps=1)
    a, b = dev(asks), dev(bids)
    for (H, W, p) in ((64, 64, 8), (128, 128, 16), (64, 32, 4)):
        st0 = venv.reset(CFG, a.clone(), b.clone(), 13)
        _, out_img = venv.step(CFG, st0, dev(blocks[0]), image_hw=(H, W), inplace=False)
        st1 = venv.reset(CFG, a.clone(), b.clone(), 13)
        _, out_pat = venv.step(CFG, st1, dev(blocks[0]), image_hw=(H, W), inplace=False, image_patch=p)
        img = out_img.image
        want = img.view(70, H // p, p, W // p, p, 2).permute(0, 1, 3, 2, 4, 5).reshape(70, (H // p) * (W // p), p * p * 2)
        assert out_pat.image.shape == want.shape and torch.equal(out_pat.image, want), (H, W, p)
        assert torch.equal(out_pat.vision_obs, out_img.vision_obs)
        if (H, W, p) == (64, 64, 8):
            cfg = vit.VIT_PARITY
            enc = vit.ViTEncoder(cfg)
            packed = vit.pack_params(cfg, vit.init_params(cfg, 0, "cuda"))
            y_img = enc.apply_packed(packed, img).clone()
            y_pat = enc.apply_packed(packed, out_pat.image, patches=True)
            assert torch.equal(y_img, y_pat)


def test_fused_ffill_with_emptied_books(c_oracle):
    """Sides that go empty mid-step exercise the forward fill and Q10 volumes."""
    E, N, M = 40, 6, 37
    rng = np.random.default_rng(5)
    asks, bids, _, msgs = fuzz_case(rng, E, N, 4, M, fill=0.3, weird=0.0)
    msgs[..., 0] = rng.choice([1, 2, 3], size=(E, M))
    msgs[..., 2] = rng.integers(20, 60, size=(E, M))                 # big cancels / aggressive orders empty the book
    ba0, bb0 = c_oracle.best_bid_ask(asks, bids)
    w = _oracle_env_step(c_oracle, asks, bids, msgs, ba0[:, 0], bb0[:, 0], 5, 0, 0)
    assert (w[3][:, :, 1] == 0).any()
    import dataclasses
    cfg = dataclasses.replace(CFG, nTradesLogged=100)
    state = venv.reset(cfg, dev(asks), dev(bids), M)
    state, out = venv.step(cfg, state, dev(msgs), n_levels=5, want_obs=True, want_raw=True)
    got = (state.ask_raw_orders, state.bid_raw_orders, state.trades, state.best_asks, state.best_bids, state.mid_price,
           out.vision_raw, out.vision_obs)
    for name, x, y in zip("asks bids trades best_asks best_bids mid raw norm".split(), w, got):
        assert host(y).tobytes() == np.ascontiguousarray(x).tobytes(), name


def test_full_size_properties():
    """BASELINE configs[1] size (4096 envs, M=13): size-independent properties, no oracle."""
    from vitmarl_b200 import synth
    E, M = 4096, 13
    l2 = synth.make_l2_books(E, 7)
    init = dev(synth.init_msgs_from_l2_batched(l2))
    a, b, t = jaxob.scan_through_entire_array(CFG, None, init, (jaxob.init_orderside(100, E), jaxob.init_orderside(100, E), None))
    assert np.array_equal(host(jaxob.get_L2_state(a, b, 10, CFG)), l2)          # init recipe round trip
    stream = synth.MessageStream(E, 7)
    vol0 = (a[..., 1].clamp(min=0).sum(1) + b[..., 1].clamp(min=0).sum(1)).long()
    msgs = dev(stream.next(M))
    (a2, b2, t2), (ba, bb) = jaxob.scan_through_entire_array_save_bidask(CFG, None, msgs, (a, b, None), M)
    # determinism / idempotence of the launch
    (a3, b3, t3), _ = jaxob.scan_through_entire_array_save_bidask(CFG, None, msgs, (a, b, None), M)
    assert torch.equal(a2, a3) and torch.equal(b2, b3) and torch.equal(t2, t3)
    # a no-op block leaves the book untouched and reports the same best prices M times
    (a4, b4, _), (ba4, _) = jaxob.scan_through_entire_array_save_bidask(CFG, None, torch.zeros_like(msgs), (a2, b2, None), M)
    assert torch.equal(a4, a2) and torch.equal(b4, b2) and torch.equal(ba4[:, 0], ba4[:, -1]) and torch.equal(ba4[:, -1], ba[:, -1])
    # book invariant: live rows have qty > 0, empty rows are all -1; books never cross
    for s in (a2, b2):
        live = s[..., 0] != -1
        assert bool((s[..., 1][live] > 0).all()) and bool((s[~live] == -1).all())
    both = (ba[:, -1, 0] != -1) & (bb[:, -1, 0] != -1)
    assert bool((ba[:, -1, 0][both] > bb[:, -1, 0][both]).all())
    # volume conservation: resting volume change = added limit qty - cancelled - 2 * traded
    traded = t2[..., 1].abs().where(t2[..., 0] != -1, torch.zeros_like(t2[..., 1])).sum(1).long()
    assert bool((traded > 0).any())


def test_fused_trade_stats_and_on_chip_trade_log(c_oracle):
    """the reward functions' trade reductions are computed inside the step kernel from the trade log while it
    is on chip; with keep_trades=False the [T,8] log is never written.  Everything else must stay byte-identical, over chained
    steps that rewrite the best-price tracks in place (the previous step's last prices are read by stride from them)."""
    from oracle import lob_oracle as O
    E, M, steps = 193, 13, 5
    asks, bids, blocks = synthetic_case(E, M, steps=steps, seed=77)
    # synthetic streams carry trader_id = order_id; make some resting / aggressing orders belong to the two agents
    ids = [-100, -37]
    rng = np.random.default_rng(3)
    for blk in blocks:
        who = rng.random((E, M))
        blk[..., 5] = np.where(who < 0.25, ids[0], np.where(who < 0.5, ids[1], blk[..., 5]))
    bufs = venv.StepBuffers()
    st_a = venv.reset(CFG, dev(asks), dev(bids), M)
    st_b = venv.reset(CFG, dev(asks), dev(bids), M)
    sentinel = st_b.trades.clone()
    ba0, bb0 = c_oracle.best_bid_ask(asks, bids)
    last_a, last_b = ba0[:, 0].copy(), bb0[:, 0].copy()
    any_trade = False
    for msgs in blocks:
        w = _oracle_env_step(c_oracle, asks, bids, msgs, last_a, last_b, 10, 0, 0)
        st_a, out_a = venv.step(CFG, st_a, dev(msgs), want_raw=True, stat_agent_ids=ids, buffers=bufs)
        st_b, out_b = venv.step(CFG, st_b, dev(msgs), want_raw=True, stat_agent_ids=ids, keep_trades=False)
        torch.cuda.synchronize()
        for st, out in ((st_a, out_a), (st_b, out_b)):
            got = (st.ask_raw_orders, st.bid_raw_orders, st.best_asks, st.best_bids, st.mid_price, out.vision_raw, out.vision_obs)
            want = (w[0], w[1], w[3], w[4], w[5], w[6], w[7])
            for name, x, y in zip("asks bids best_asks best_bids mid raw norm".split(), want, got):
                assert host(y).tobytes() == np.ascontiguousarray(x).tobytes(), name
        assert host(st_a.trades).tobytes() == w[2].tobytes()
        assert torch.equal(st_b.trades, sentinel)                       # never written
        want_stats = np.stack([np.stack([O.agent_trade_stats(w[2][e], a, CFG.tick_size) for a in ids]) for e in range(E)])
        assert np.array_equal(host(out_a.trade_stats), want_stats) and np.array_equal(host(out_b.trade_stats), want_stats)
        any_trade = any_trade or bool((want_stats[..., 1] > 0).any())
        asks, bids, last_a, last_b = w[0], w[1], w[3][:, -1, 0].copy(), w[4][:, -1, 0].copy()
    assert any_trade
    with pytest.raises(_capi.VitmarlError):
        venv.step(CFG, st_b, dev(blocks[0]), keep_trades=False)          # the log may stay on chip only when its reductions are asked for


def test_world_clock_in_the_fused_step():
    """marl_env.py:406,468,482: time <- the (s, ns) columns of the step's last message, delta_time in float32 with one rounding
    per operation -- bit-exact vs the NumPy float32 restatement, over chained steps (the clock is updated in place), with
    times around the LOBSTER session (34 200 .. 57 600 s, where float32 resolves ~4 ms) and at the int32 extremes."""
    import dataclasses
    from oracle import lob_oracle as O
    E, M, steps = 97, 7, 4
    asks, bids, blocks = synthetic_case(E, M, steps=steps, seed=5)
    rng = np.random.default_rng(11)
    t0 = np.stack([rng.integers(34200, 57600, E), rng.integers(0, 1_000_000_000, E)], 1).astype(np.int32)
    t0[0] = (0, 0); t0[1] = (2**31 - 1, 999_999_999); t0[2] = (57599, 1)
    st = dataclasses.replace(venv.reset(CFG, dev(asks), dev(bids), M), time=dev(t0))
    tref = t0.copy()
    for k, msgs in enumerate(blocks):
        msgs = msgs.copy()
        msgs[:, -1, 6] = tref[:, 0] + rng.integers(0, 3, E)                      # the last message carries the step's final time
        msgs[:, -1, 7] = rng.integers(0, 1_000_000_000, E)
        if k == 1:
            msgs[3, -1, 6:8] = (2**31 - 1, 0); msgs[4, -1, 6:8] = (0, 5)
        st, _ = venv.step(CFG, st, dev(msgs), want_obs=False)
        torch.cuda.synchronize()
        want = [O.world_time_update(msgs[e], tref[e]) for e in range(E)]
        wt = np.stack([w[0] for w in want]); wd = np.array([w[1] for w in want], np.float32)
        assert np.array_equal(host(st.time), wt)
        assert host(st.delta_time).tobytes() == wd.tobytes()
        tref = wt
    # an untracked clock stays untracked; half a clock is rejected by the ABI
    st2, _ = venv.step(CFG, venv.reset(CFG, dev(asks), dev(bids), M), dev(blocks[0]), want_obs=False)
    assert st2.time is None and st2.delta_time is None
    import ctypes
    a = _capi.EnvStepArgs()
    a.E, a.N, a.T, a.M = 1, 100, 100, 1
    a.time_in = 16                                                               # (never dereferenced: the argument check comes first)
    assert _capi.lib().vitmarl_env_step2(None, ctypes.byref(a)) == _capi.EINVAL
