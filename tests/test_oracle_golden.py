"""Pins the oracle (NumPy and C restatements) against the reference's own known-answer
material: G1 (notebook-stored LOBSTER rows) and G2 (jorderbook.py worked example), plus a
quirk-by-quirk suite for SURVEY.md section 8a Q1-Q13 with hand-derived expectations."""
import numpy as np
import pytest

from oracle import lob_oracle as O
from tests.util import lobster_to_msg


def _empty(N=100, T=100):
    return O.init_orderside(N), O.init_orderside(N), np.full((T, 8), -1, np.int32)


def _run_c(C, msgs, a, b, t):
    a2, b2, t2, ba, bb = C.lob_step(a[None], b[None], np.asarray(msgs, np.int32)[None], trades_in=t[None])
    return a2[0], b2[0], t2[0], ba[0], bb[0]


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_g1_lobster_rows(golden, c_oracle, impl):
    rows = np.array(golden["book_rows"], dtype=np.int32)
    init = O.init_msgs_from_l2(rows[0], (34200, 17459617))
    a, b, t = _empty()
    if impl == "numpy":
        a, b, t = O.scan_through_entire_array(init, (a, b, t))
    else:
        a, b, t, _, _ = _run_c(c_oracle, init, a, b, t)
    assert np.array_equal(O.get_L2_state(a, b, 10), rows[0])
    for i in range(1, 5):
        m = lobster_to_msg(golden["messages_lobster"][i])[None]
        if impl == "numpy":
            a, b, t = O.scan_through_entire_array(m, (a, b, t))
            l2 = O.get_L2_state(a, b, 10)
        else:
            a, b, t, _, _ = _run_c(c_oracle, m, a, b, t)
            l2 = O.get_L2_state(a, b, 10)
            raw, _, _ = c_oracle.render(a[None], b[None], n_levels=10)
            assert np.array_equal(raw[0], O.get_vision_L2_state(a, b, 10))
        assert np.array_equal(l2, rows[i]), f"row {i}"
    assert (t == -1).all()          # four passive limit orders: no trades


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_g2_worked_example(golden, c_oracle, impl):
    g2 = golden["g2"]
    init = O.init_msgs_from_l2(np.array(g2["l2init"]), (0, 0))
    msgs = np.array(g2["msgs"], dtype=np.int32)
    a, b, t = _empty()
    if impl == "numpy":
        a, b, t = O.scan_through_entire_array(init, (a, b, t))
        (a, b, t), (ba, bb) = O.scan_through_entire_array_save_bidask(msgs, (a, b, t), 2)
    else:
        a, b, t, _, _ = _run_c(c_oracle, init, a, b, t)
        a, b, t, ba, bb = _run_c(c_oracle, msgs, a, b, t)
    # second message (sell 2 @ 346000) trades 2 @ 350100 against the best bid (89 -> 87)
    assert t[0].tolist() == [350100, 2, -2, 8777, 3401, 5060000, -3, 8777]
    assert (t[1:] == -1).all()
    assert ba.tolist() == [[354200, 452], [354200, 452]]
    assert bb.tolist() == [[350100, 89], [350100, 87]]
    # first message rests 99 @ 346000 as the second bid level
    l2 = O.get_L2_state(a, b, 2)
    assert l2.tolist() == [354200, 452, 350100, 87, 361200, 100, 346000, 99]


def _msg(t, s, q, p, oid=50, tid=60, ts=7, tns=8):
    return np.array([t, s, q, p, oid, tid, ts, tns], dtype=np.int32)


def _both(c_oracle, msgs, a, b, t):
    msgs = np.asarray(msgs, np.int32).reshape(-1, 8)
    (a1, b1, t1), (ba1, bb1) = O.scan_through_entire_array_save_bidask(msgs, (a, b, t), len(msgs))
    a2, b2, t2, ba2, bb2 = _run_c(c_oracle, msgs, a, b, t)
    for x, y in ((a1, a2), (b1, b2), (t1, t2), (ba1, ba2), (bb1, bb2)):
        assert np.array_equal(x, y)
    return a1, b1, t1, ba1, bb1


def test_q1_partial_minus_one_row_is_empty(c_oracle):
    a, b, t = _empty(4, 4)
    a[0] = [100, 5, 1, 1, 1, 1]
    a[1] = [101, 5, 2, -1, 1, 1]            # trader id -1 -> counts as an empty row (Q1)
    a2, *_ = _both(c_oracle, [_msg(1, -1, 3, 105)], a, b, t)
    assert a2[1].tolist() == [105, 3, 50, 60, 7, 8]


def test_q2_full_side_overwrites_last_row(c_oracle):
    a, b, t = _empty(3, 4)
    for i in range(3):
        a[i] = [100 + i, 5, i + 1, 1, 1, 1]
    a2, *_ = _both(c_oracle, [_msg(1, -1, 3, 110)], a, b, t)
    assert a2[2].tolist() == [110, 3, 50, 60, 7, 8] and a2[0].tolist() == [100, 5, 1, 1, 1, 1]
    # even a fully matched (zero remainder) order overwrites, then the row is wiped
    b[0] = [120, 9, 77, 1, 1, 1]
    a3, b3, t3, *_ = _both(c_oracle, [_msg(1, -1, 9, 115)], a, b, t)
    assert (a3[2] == -1).all() and (b3[0] == -1).all() and t3[0, :2].tolist() == [120, 9]


def test_q3_nonpositive_quantity_rows_are_wiped(c_oracle):
    a, b, t = _empty(4, 4)
    a[0] = [100, 0, 1, 1, 1, 1]             # stale zero-qty row wiped by the next op on that side
    a2, *_ = _both(c_oracle, [_msg(1, -1, -5, 101)], a, b, t)   # negative qty -> max(0,q)=0 -> wiped
    assert (a2 == -1).all()


def test_q4_cancel_falls_back_to_init_orders(c_oracle):
    a, b, t = _empty(4, 4)
    b[0] = [100, 50, -2, -2, 1, 1]
    b[1] = [100, 50, -200, -3, 1, 1]        # agent ids count down from -200: also <= init_id
    b[2] = [100, 5, 9, 9, 1, 1]
    _, b2, *_ = _both(c_oracle, [_msg(2, 1, 20, 100, oid=12345)], a, b, t)
    assert b2[0, 1] == 30
    _, b3, *_ = _both(c_oracle, [_msg(3, 1, 40, 100, oid=12345)], a, b2, t)   # qty 40 > 30 -> second row
    assert b3[0, 1] == 30 and b3[1, 1] == 10


def test_q5_cancel_without_match_hits_last_row(c_oracle):
    a, b, t = _empty(3, 4)
    a[2] = [105, 10, 4, 4, 1, 1]
    a2, *_ = _both(c_oracle, [_msg(2, -1, 3, 999, oid=777)], a, b, t)
    assert a2[2, 1] == 7
    a3, *_ = _both(c_oracle, [_msg(2, -1, 7, 999, oid=777)], a2, b, t)
    assert (a3[2] == -1).all()


def test_q6_q7_trade_slot_and_sign(c_oracle):
    a, b, t = _empty(4, 2)
    b[0] = [100, 5, 1, 11, 1, 1]
    b[1] = [100, 5, 2, 12, 1, 2]
    b[2] = [99, 5, 3, 13, 1, 1]
    # type 4 with side +1 dispatches to ask_lim but keeps side=+1 in the trade sign (Q7)
    _, b2, t2, *_ = _both(c_oracle, [_msg(4, 1, 12, 99)], a, b, t)
    assert t2[0].tolist() == [100, -5, 1, 50, 7, 8, 11, 60]
    # trades full (T=2): third fill overwrites the last row (Q6)
    assert t2[1].tolist() == [99, -2, 3, 50, 7, 8, 13, 60]
    assert b2[2, 1] == 3
    # a trade whose time_s is -1 leaves its slot "free" and is overwritten by the next one
    a, b, t = _empty(4, 4)
    b[0] = [100, 5, 1, 11, 1, 1]
    b[1] = [99, 5, 2, 12, 1, 1]
    _, _, t3, *_ = _both(c_oracle, [_msg(1, -1, 10, 99, ts=-1)], a, b, t)
    assert t3[0, 0] == 99 and (t3[1:] == -1).all()


def test_q8_unknown_type_side_goes_to_ask_lim(c_oracle):
    a, b, t = _empty(4, 4)
    a2, *_ = _both(c_oracle, [_msg(5, 0, 3, 105), _msg(0, 0, 0, 0, 0, 0, 0, 0)], a, b, t)
    assert a2[0].tolist() == [105, 3, 50, 60, 7, 8] and (a2[1:] == -1).all()


def test_q9_q10_empty_sides(c_oracle):
    a, b, t = _empty(5, 4)
    *_, ba, bb = _both(c_oracle, [_msg(0, 0, 0, 0, 0, 0, 0, 0)], a, b, t)
    assert ba.tolist() == [[-1, -5]] and bb.tolist() == [[-1, -5]]      # Q10: -(#empty rows)
    a2, b2, t2, *_ = _both(c_oracle, [_msg(1, 1, 5, 2_000_000)], a, b, t)   # nothing to match (Q9)
    assert (t2 == -1).all() and b2[0, 0] == 2_000_000


def test_q11_negative_remainder_not_rested(c_oracle):
    a, b, t = _empty(4, 4)
    a[0] = [100, 50, 1, 1, 1, 1]
    a2, b2, t2, *_ = _both(c_oracle, [_msg(1, 1, 20, 100)], a, b, t)
    assert a2[0, 1] == 30 and (b2 == -1).all() and t2[0, 1] == -20


def test_price_time_priority_and_tie_break(c_oracle):
    a, b, t = _empty(5, 5)
    a[0] = [100, 5, 1, 1, 10, 5]
    a[1] = [100, 5, 2, 2, 9, 9]      # same price, earlier second -> first
    a[2] = [100, 5, 3, 3, 9, 3]      # same second, earlier ns -> very first
    a[3] = [100, 5, 4, 4, 9, 3]      # identical time -> lower row index wins
    _, _, t2, *_ = _both(c_oracle, [_msg(1, 1, 20, 100)], a, b, t)
    assert t2[:4, 2].tolist() == [3, 4, 2, 1]


def test_q12_vision_levels_and_padding(c_oracle):
    a, b, t = _empty(6, 4)
    a[0] = [105, 5, 1, 1, 1, 1]
    a[1] = [103, 2, 2, 1, 1, 1]
    a[2] = [105, 7, 3, 1, 1, 1]
    b[0] = [99, 4, 4, 1, 1, 1]
    v = O.get_vision_L2_state(a, b, 4)
    assert v[:, 0, 0].tolist() == [103, 105, -1, -1] and v[:, 1, 0].tolist() == [2, 12, 0, 0]
    assert v[:, 0, 1].tolist() == [99, -1, -1, -1] and v[:, 1, 1].tolist() == [4, 0, 0, 0]
    raw, _, _ = c_oracle.render(a[None], b[None], n_levels=4)
    assert np.array_equal(raw[0], v)
    assert O.get_L2_state(a, b, 2).tolist() == [103, 2, 99, 4, 105, 12, -1, 0]


def test_q13_init_messages():
    m = O.init_msgs_from_l2(np.arange(1, 9), (5, 6))
    assert m[:, 0].tolist() == [1, 1, 1, 1] and m[:, 1].tolist() == [-1, 1, -1, 1]
    assert m[:, 3].tolist() == [1, 3, 5, 7] and m[:, 2].tolist() == [2, 4, 6, 8]
    assert m[:, 4].tolist() == [-2] * 4 and m[:, 5].tolist() == [-2, -3, -4, -5]
    assert m[:, 6].tolist() == [5] * 4 and m[:, 7].tolist() == [6] * 4


def test_ffill_and_mid(c_oracle):
    pa = np.array([[-1, -3], [-1, -3], [105, 4], [-1, -3]], np.int32)
    pb = np.array([[99, 2], [-1, -3], [98, 1], [98, 1]], np.int32)
    fa, fb = O.ffill_best_prices(pa, 104), O.ffill_best_prices(pb, 97)
    assert fa.tolist() == [[104, 0], [104, 0], [105, 4], [105, 0]]
    assert fb.tolist() == [[99, 2], [99, 0], [98, 1], [98, 1]]
    ca, cb, mid = c_oracle.ffill_mid(pa[None], pb[None], np.array([104]), np.array([97]))
    assert np.array_equal(ca[0], fa) and np.array_equal(cb[0], fb)
    assert mid[0] == np.float32(101.5) == O.mid_price_f32(98, 105)
    assert O.ffill_best_prices(pa, -1)[:2].tolist() == [[-1, 0], [-1, 0]]


def test_normalize_vision_obs_values():
    raw = np.array([[[1010, 990], [9, 4]], [[1030, -1], [0, 7]], [[-1, -1], [5, 5]]], np.int32)
    out = O.normalize_vision_obs(raw, np.float32(1000.0), 10)
    assert out.dtype == np.float32 and out.shape == (3, 3, 2)
    np.testing.assert_allclose(out[:, 0, 0], [1.0, 3.0, 0.0])
    np.testing.assert_allclose(out[:, 0, 1], [1.0, 0.0, 0.0])
    np.testing.assert_allclose(out[:, 1, 0], np.log1p([9.0, 0.0, 0.0]).astype(np.float32), rtol=1e-7)
    np.testing.assert_allclose(out[:, 2, 0], np.log1p([9.0, 9.0, 0.0]).astype(np.float32), rtol=1e-7)
    np.testing.assert_allclose(out[:, 1, 1], np.log1p([4.0, 0.0, 0.0]).astype(np.float32), rtol=1e-7)


def test_log1p_is_correctly_rounded_on_small_integers():
    import math
    bad = 0
    for v in list(range(0, 20000)) + [2 ** k for k in range(15, 31)] + [2 ** 31 - 1]:
        x = np.float32(v)
        got, want = O.log1p_f32(x), np.float32(math.log1p(float(x)))
        bad += int(got.tobytes() != want.tobytes())
    assert bad == 0


def test_render_image_spec():
    assert [O.bar_length(v, 64) for v in (0, 1, 2, 3, 100, 5000, 65534, 65535, 10 ** 9)] == [0, 4, 6, 8, 26, 48, 63, 64, 64]
    a, b, _ = _empty(6, 4)
    a[0] = [1000, 3, 1, 1, 1, 1]
    a[1] = [1250, 100, 2, 1, 1, 1]     # (1250-1000)//100 = row 2
    a[2] = [1299, 1, 3, 1, 1, 1]       # same row 2
    b[0] = [900, 1, 4, 1, 1, 1]
    b[1] = [100, 50, 5, 1, 1, 1]       # row 8 -> outside H=8
    img = O.render_image(a, b, 8, 64, 100)
    assert img[:, :, 0].sum(axis=1).tolist() == [8, 0, O.bar_length(101, 64), 0, 0, 0, 0, 0]
    assert img[:, :, 1].sum(axis=1).tolist() == [4, 0, 0, 0, 0, 0, 0, 0]
    assert (img[2, :26, 0] == 1).all() and (img[2, 26:, 0] == 0).all()


def test_world_clock_hand_derived():
    """marl_env.py:406,468: final_time = last message's (s, ns); delta in float32.  34201.5 - 34200 - 0.25 = 1.25 exactly; at
    57 599 s a nanosecond is far below the float32 grid (2^-8 s), so the delta collapses onto it."""
    from oracle import lob_oracle as O
    msgs = np.zeros((3, 8), np.int32)
    msgs[-1, 6:8] = (34201, 500_000_000)
    ft, d = O.world_time_update(msgs, np.array([34200, 250_000_000], np.int32))
    assert ft.tolist() == [34201, 500_000_000] and d == np.float32(1.25) and d.dtype == np.float32
    msgs[-1, 6:8] = (57599, 1)
    _, d = O.world_time_update(msgs, np.array([57599, 0], np.int32))
    assert d == np.float32(0.0)
