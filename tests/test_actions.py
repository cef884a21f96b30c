"""the default action -> message tables of the two agents.  CPU: the NumPy restatement (oracle/action_oracle.py)
against hand-derived rows.  GPU: CUDA == NumPy bit for bit over random and adversarial inputs (prices beyond float32's integer
range, negative / empty prices, odd spreads, every action, quantity overflow of the task)."""
import numpy as np
import pytest

from oracle import action_oracle as A


def test_exec_fixed_quants_complex_hand_derived():
    # buy task, action 5 = far touch x 2: best ask 31 200 150 -> 31 200 100, best bid 31 199 950 -> 31 199 900
    m = A.exec_action_msgs_fixed_quants_complex(5, 31_200_150, 31_199_950, 0, 500, 100, (34200, 5), 1000001)
    assert m.tolist() == [[1, 1, 20, 31_200_100, -9, 1000001, 34200, 5], [1, 1, 0, 31_200_000, -9, 1000001, 34200, 5],
                          [1, 1, 0, 31_199_900, -9, 1000001, 34200, 5], [1, 1, 0, 31_199_800, -9, 1000001, 34200, 5]]
    # sell task, action 6 = mid x 2 = 20 > 5 left: everything that is left goes to the first message (vision_env.py:1128-1132)
    m = A.exec_action_msgs_fixed_quants_complex(6, 31_200_150, 31_199_950, 1, 500, 495, (34200, 5), 7, time_delay_obs_act=3)
    assert m[:, 2].tolist() == [5, 0, 0, 0] and m[:, 1].tolist() == [-1] * 4 and m[0, 6:].tolist() == [34203, 8]
    assert m[:, 3].tolist() == [31_199_900, 31_200_000, 31_200_100, 31_200_200]
    # the sell-side mid goes through float32: 67 108 900 + 67 109 000 is not representable, the buy-side integer path differs
    b = A.exec_action_msgs_fixed_quants_complex(2, 67_109_000, 67_108_900, 0, 500, 0, (0, 0), 1)
    s = A.exec_action_msgs_fixed_quants_complex(2, 67_109_000, 67_108_900, 1, 500, 0, (0, 0), 1)
    assert b[1, 3] == 67_108_900 and s[1, 3] == int(np.float32(np.ceil(A.float_floor_divide(np.float32(np.float32(134_217_900) / np.float32(2)), 100)) * np.float32(100)))


def test_mm_spread_skew_hand_derived():
    # action 3 = wide spread (x3), bid skew (-5 ticks): spread 200 -> 600, mid 31 200 000 - 500, half spread 300
    m = A.mm_action_msgs_spread_skew(3, 31_200_150, 31_199_950, (34200, 5), 7)
    assert m.tolist() == [[1, 1, 10, 31_199_200, -9, 7, 34200, 5], [1, -1, 10, 31_199_800, -9, 7, 34200, 5]]
    # action 1 = tight, neutral: quotes at mid -/+ half the current spread, floored to the tick
    m = A.mm_action_msgs_spread_skew(1, 31_200_100, 31_199_800, (1, 2), 7)
    assert m[:, 3].tolist() == [31_199_800, 31_200_100]
    assert A.float_floor_divide(-250.0, 100) == -3.0 and A.float_floor_divide(250.0, -100) == -3.0 and A.float_floor_divide(300.0, 100) == 3.0


@pytest.mark.gpu
def test_cuda_matches_numpy_bit_for_bit():
    import torch
    from vitmarl_b200 import actions
    rng = np.random.default_rng(0)
    E, M = 3000, 3
    mid = rng.integers(1_000, 90_000_000, E)
    spread = rng.integers(0, 5_000, E)
    ba = (mid + spread).astype(np.int64); bb = (mid - rng.integers(0, 5_000, E)).astype(np.int64)
    ba[:8] = [-1, 0, 99, 2**31 - 1, 67_109_000, 100, 2**31 - 1, 31_200_150]
    bb[:8] = [-1, -1, -101, 2**31 - 1, 67_108_900, 100, -2**31, 31_199_950]
    best_asks = np.zeros((E, M, 2), np.int32); best_bids = np.zeros((E, M, 2), np.int32)
    best_asks[:, -1, 0] = ba.astype(np.int32); best_bids[:, -1, 0] = bb.astype(np.int32)
    best_asks[:, :-1] = 12345; best_bids[:, :-1] = 54321                       # only [-1][0] may be read
    time = np.stack([rng.integers(0, 2**31 - 1, E), rng.integers(0, 10**9, E)], 1).astype(np.int32)
    sell = rng.integers(0, 2, E).astype(np.int32)
    task = rng.integers(0, 1000, E).astype(np.int32); done = rng.integers(0, 1000, E).astype(np.int32)
    task[:4] = [2**31 - 1, 5, 0, 20]; done[:4] = [-5, 0, 0, 0]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for delay, fq, nt in ((0, 10, 1), (7, 3, 2)):
        act = rng.integers(0, 13, E).astype(np.int32)
        got = actions.getActionMsgs_fixedQuant_complex(cu(act), cu(best_asks), cu(best_bids), cu(sell), cu(task), cu(done), cu(time), 1000001,
                                                       n_ticks_in_book=nt, fixed_quant_value=fq, time_delay_obs_act=delay).cpu().numpy()
        want = np.stack([A.exec_action_msgs_fixed_quants_complex(int(act[e]), int(best_asks[e, -1, 0]), int(best_bids[e, -1, 0]), int(sell[e]),
                                                                 int(task[e]), int(done[e]), time[e], 1000001, n_ticks_in_book=nt,
                                                                 fixed_quant_value=fq, time_delay_obs_act=delay) for e in range(E)])
        assert np.array_equal(got, want)
    for mt, sm, km in (("tick", 3.0, 5.0), ("spread", 2.5, 0.75), ("tick", 1.0, 100.0)):
        act = rng.integers(0, 6, E).astype(np.int32)
        got = actions.getActionMsgs_spread_skew(cu(act), cu(best_asks), cu(best_bids), cu(time), 7, spread_multiplier=sm, skew_multiplier=km,
                                                multiplier_type=mt, time_delay_obs_act=2).cpu().numpy()
        want = np.stack([A.mm_action_msgs_spread_skew(int(act[e]), int(best_asks[e, -1, 0]), int(best_bids[e, -1, 0]), time[e], 7,
                                                      spread_multiplier=sm, skew_multiplier=km, multiplier_type=mt, time_delay_obs_act=2)
                         for e in range(E)])
        assert np.array_equal(got, want)
