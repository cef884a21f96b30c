"""Callers either side of the book step: oracle known answers on CPU, CUDA parity on GPU."""
import numpy as np
import pytest

from oracle import lob_oracle as O


def _book(rng, N=12):
    b = np.full((N, 6), -1, np.int32)
    live = rng.random(N) < 0.7
    b[live] = np.stack([rng.integers(100, 110, N), rng.integers(1, 50, N), rng.integers(1, 99, N), rng.integers(-3, 2, N),
                        rng.integers(0, 5, N), rng.integers(0, 5, N)], axis=1)[live]
    return b


def test_get_cancel_msgs_known_answer():
    b = np.full((5, 6), -1, np.int32)
    b[1] = [101, 7, 11, -100, 3, 4]
    b[3] = [102, 9, 12, -100, 3, 5]
    b[4] = [103, 1, 13, -101, 3, 6]
    out = O.getCancelMsgs(b, -100, 3, 1, 77, 88)
    assert out.tolist() == [[2, 1, 7, 101, 11, -100, 77, 88], [2, 1, 9, 102, 12, -100, 77, 88], [2, 1, 0, 0, 0, 0, 77, 88]]
    assert O.getCancelMsgs(b, -100, 1, -1, 1, 2).tolist() == [[2, -1, 7, 101, 11, -100, 1, 2]]


def test_filter_messages_known_answers():
    """The docstring example of the reference (at one level, 3 cancels & 1 action -> 2 cancels, 0 actions), the 'new order
    larger than the old one' case (nothing is netted), and the pairing quirk: the t-th matching action is netted against
    the t-th matching cancellation BY POSITION, whatever their prices."""
    def rows(kind, specs):
        m = np.zeros((4, 8), np.int32)
        for i, (q, p) in enumerate(specs):
            m[i] = [kind, 1, q, p, -150 - i, -100, 34200, 0]
        return m
    ao, co = O.filter_messages(rows(1, [(10, 2200100)]), rows(2, [(10, 2200100)] * 3))
    assert (ao == 0).all() and co[:, 2].tolist() == [0, 10, 10, 0]
    ao, co = O.filter_messages(rows(1, [(10, 2200100)]), rows(2, [(5, 2200100)]))
    assert ao[0, 2] == 10 and co[0, 2] == 5                      # 10 > 5: the old order is cancelled in full, the new one placed in full
    a = rows(1, [(10, 2200100), (10, 2200300)])
    c = rows(2, [(10, 2200100), (10, 2200100), (10, 2200100), (5, 2200300)])
    ao, co = O.filter_messages(a, c)
    assert (ao == 0).all() and co[:, 2].tolist() == [0, 0, 10, 5]   # action 1 @ 2200300 was paired with cancel 1 @ 2200100
    # price 0 never matches; the inputs are not modified
    z = np.zeros((2, 8), np.int32); z[:, 2] = 7
    ao, co = O.filter_messages(z, z)
    assert np.array_equal(ao, z) and np.array_equal(co, z) and (z[:, 2] == 7).all()


def test_agent_trade_stats_known_answer():
    """Two agent trades (one passive buy of 5 @ 2 200 100, one aggressive sell of 3 @ 2 200 300), one foreign trade, one empty row."""
    t = np.full((4, 8), -1, np.int32)
    t[0] = [2200100, 5, 11, 12, 34200, 1, -100, 7]       # qty >= 0 and agent passive -> agent BUYS 5
    t[1] = [2200300, 3, 13, 14, 34200, 2, 8, -100]       # qty >= 0 and agent aggressive -> agent SELLS 3
    t[2] = [2200200, -4, 15, 16, 34200, 3, 8, 9]         # others
    s = O.agent_trade_stats(t, -100, 100)
    assert s.tolist() == [8, 8, 22001 * 5 + 22003 * 3, 5, 3, 8, 2, 4]
    assert O.agent_trade_stats(t, -1, 100).tolist() == [0, 0, 0, 0, 0, 0, 0, 12]   # the empty row never counts, even for id -1
    assert np.array_equal(O.get_agent_trades(t, -100)[:, 1], [5, 3, 0, 0])


def test_get_agent_trades_known_answer():
    t = np.full((4, 8), -1, np.int32)
    t[0] = [100, 5, 1, 2, 3, 4, -100, 9]
    t[1] = [100, -5, 1, 2, 3, 4, 8, -100]
    t[2] = [100, 5, 1, 2, 3, 4, 8, 9]
    out = O.get_agent_trades(t, -100)
    assert out[0].tolist() == t[0].tolist() and out[1].tolist() == t[1].tolist() and (out[2:] == 0).all()
    assert (O.get_agent_trades(t, -1) == 0).all()      # empty rows (price -1) never count, even for id -1


def test_build_step_msgs_known_answer():
    md = np.arange(10 * 8, dtype=np.int32).reshape(10, 8)
    md[:, 6] = np.arange(10) + 100
    cancel = np.full((2, 8), 7, np.int32)
    action = np.stack([np.full(8, 10 + k, np.int32) for k in range(3)])
    comb, newc = O.build_step_msgs(md, 2, 1, 2, cancel, action, -200, perm=[2, 0, 1], end_time_s=105)
    assert comb.shape == (7, 8) and newc == -203
    assert comb[2:5, 4].tolist() == [-202, -200, -201] and comb[2:5, 0].tolist() == [12, 10, 11]
    assert np.array_equal(comb[5], md[4])                       # time 104 < 105 kept
    assert comb[6, :6].tolist() == [0] * 6 and comb[6, 6:].tolist() == md[5, 6:].tolist()   # time 105 >= end -> zeroed
    # dynamic_slice clamps the start so that the slice fits
    assert np.array_equal(O.get_data_messages(md, 9, 5, 3), md[7:10])


def test_auto_reset_known_answer():
    init_a = np.arange(2 * 3 * 6, dtype=np.int32).reshape(2, 3, 6)
    init_b = -init_a
    asks = np.full((3, 3, 6), 7, np.int32); bids = np.full((3, 3, 6), 8, np.int32)
    trades = np.full((3, 2, 8), 9, np.int32)
    ba = np.full((3, 4, 2), 1, np.int32); bb = np.full((3, 4, 2), 2, np.int32)
    mid = np.array([1.5, 2.5, 3.5], np.float32)
    out = O.auto_reset([0, 1, 0], [0, 1, 0], init_a, init_b, [[101, 5], [2200100, 7]], [[99, 6], [2199900, 8]], asks, bids, trades, ba, bb, mid)
    a, b, t, oa, ob, m = out
    assert (a[0] == 7).all() and (a[2] == 7).all() and np.array_equal(a[1], init_a[1]) and np.array_equal(b[1], init_b[1])
    assert (t[1] == -1).all() and (t[0] == 9).all()
    assert oa[1].tolist() == [[2200100, 7]] * 4 and ob[1].tolist() == [[2199900, 8]] * 4 and (oa[0] == 1).all()
    assert m.tolist() == [1.5, 2200000.0, 3.5]


@pytest.mark.gpu
def test_auto_reset_cuda_parity():
    import torch
    from vitmarl_b200 import env as venv
    rng = np.random.default_rng(11)
    E, N, T, M, nW = 133, 100, 100, 13, 9
    init_a = rng.integers(-1, 3_000_000, (nW, N, 6)).astype(np.int32); init_b = rng.integers(-1, 3_000_000, (nW, N, 6)).astype(np.int32)
    init_t = rng.integers(-1, 50, (nW, T, 8)).astype(np.int32)
    iba = rng.integers(-1, 2**31 - 1, (nW, 2)).astype(np.int32); ibb = rng.integers(-1, 2**31 - 1, (nW, 2)).astype(np.int32)
    for use_trades in (True, False):
        asks = rng.integers(-1, 99, (E, N, 6)).astype(np.int32); bids = rng.integers(-1, 99, (E, N, 6)).astype(np.int32)
        trades = rng.integers(-1, 99, (E, T, 8)).astype(np.int32)
        ba = rng.integers(-1, 99, (E, M, 2)).astype(np.int32); bb = rng.integers(-1, 99, (E, M, 2)).astype(np.int32)
        mid = rng.random(E).astype(np.float32)
        done = (rng.random(E) < 0.4).astype(np.int32); win = rng.integers(0, nW, E).astype(np.int32)
        want = O.auto_reset(done, win, init_a, init_b, iba, ibb, asks, bids, trades, ba, bb, mid, init_t if use_trades else None)
        dev = lambda x: torch.from_numpy(x).cuda()
        st = venv.BookState(dev(asks), dev(bids), dev(trades), dev(ba), dev(bb), dev(mid))
        venv.auto_reset(st, dev(done), dev(win), dev(init_a), dev(init_b), dev(iba), dev(ibb), dev(init_t) if use_trades else None)
        got = (st.ask_raw_orders, st.bid_raw_orders, st.trades, st.best_asks, st.best_bids, st.mid_price)
        for g, w in zip(got, want):
            assert g.cpu().numpy().tobytes() == np.ascontiguousarray(w).tobytes()


@pytest.mark.gpu
def test_env_glue_cuda_parity():
    import torch
    from vitmarl_b200 import env as venv, jaxob
    rng = np.random.default_rng(3)
    E, N, T = 70, 37, 20
    books = np.stack([_book(rng, N) for _ in range(E)])
    ct = rng.integers(0, 1000, (E, 2)).astype(np.int32)
    for agent, size, side in ((-3, 4, 1), (0, 40, -1), (1, 1, 1)):
        got = jaxob.getCancelMsgs(torch.from_numpy(books).cuda(), agent, size, side, torch.from_numpy(ct).cuda()).cpu().numpy()
        want = np.stack([O.getCancelMsgs(books[e], agent, size, side, ct[e, 0], ct[e, 1]) for e in range(E)])
        assert np.array_equal(got, want)
    trades = rng.integers(-3, 4, (E, T, 8)).astype(np.int32)
    for agent in (-1, 0, 2):
        got = jaxob.get_agent_trades(torch.from_numpy(trades).cuda(), agent).cpu().numpy()
        want = np.stack([O.get_agent_trades(trades[e], agent) for e in range(E)])
        assert np.array_equal(got, want)
    for n in (1, 4, 8, 13):
        am = rng.integers(-2, 4, (E, n, 8)).astype(np.int32)
        cm = rng.integers(-2, 4, (E, n, 8)).astype(np.int32)
        am[..., 3] = rng.integers(0, 4, (E, n)); cm[..., 3] = rng.integers(0, 4, (E, n))      # few price levels: many matches
        ga, gc = jaxob.filter_messages(torch.from_numpy(am).cuda(), torch.from_numpy(cm).cuda())
        for e in range(E):
            wa, wc = O.filter_messages(am[e], cm[e])
            assert np.array_equal(ga[e].cpu().numpy(), wa) and np.array_equal(gc[e].cpu().numpy(), wc), (n, e)
    big = trades.copy()
    big[:, :, 0] = rng.integers(-1, 2_300_000, (E, T))
    big[:, :, 1] = rng.integers(-2**31, 2**31 - 1, (E, T), dtype=np.int64).astype(np.int32) * (rng.random((E, T)) < 0.1) + rng.integers(-500, 500, (E, T))
    for tr_, tick in ((trades, 1), (big, 100)):
        for agent in (-1, 0, 2):
            got = jaxob.agent_trade_stats(torch.from_numpy(tr_).cuda(), agent, tick).cpu().numpy()
            want = np.stack([O.agent_trade_stats(tr_[e], agent, tick) for e in range(E)])
            assert np.array_equal(got, want)
    n_total, n_data, Mc, Ma = 500, 7, 5, 12
    md = rng.integers(-5, 50000, (n_total, 8)).astype(np.int32)
    md[:, 6] = np.sort(rng.integers(34200, 34300, n_total))
    start = rng.integers(-3, n_total + 5, E).astype(np.int32)
    stepc = rng.integers(0, 20, E).astype(np.int32)
    cancel = rng.integers(-9, 9, (E, Mc, 8)).astype(np.int32)
    action = rng.integers(-9, 9, (E, Ma, 8)).astype(np.int32)
    counter = rng.integers(-5000, -200, E).astype(np.int32)
    perm = np.stack([rng.permutation(Ma) for _ in range(E)]).astype(np.int32)
    endt = rng.integers(34200, 34300, E).astype(np.int32)
    dev = lambda a: torch.from_numpy(a).cuda()
    for use_perm, use_end in ((True, True), (False, False)):
        got, newc = venv.build_step_msgs(dev(md), dev(start), dev(stepc), n_data, dev(cancel), dev(action), dev(counter),
                                         dev(perm) if use_perm else None, dev(endt) if use_end else None)
        for e in range(E):
            w, wc = O.build_step_msgs(md, start[e], stepc[e], n_data, cancel[e], action[e], counter[e],
                                      perm[e] if use_perm else None, endt[e] if use_end else None)
            assert np.array_equal(got[e].cpu().numpy(), w) and int(newc[e]) == wc
