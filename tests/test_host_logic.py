"""Host-side logic (CPU): synthetic LOBSTER-format generator and the batched A8 init recipe."""
import numpy as np

from oracle import lob_oracle as O
from vitmarl_b200 import synth


def test_l2_books_are_valid_lobster_rows():
    l2 = synth.make_l2_books(32, seed=3)
    assert l2.shape == (32, 40) and l2.dtype == np.int32
    ask_p, bid_p = l2[:, 0::4], l2[:, 2::4]
    assert (np.diff(ask_p, axis=1) > 0).all() and (np.diff(bid_p, axis=1) < 0).all()
    assert (ask_p[:, 0] == 2_200_100).all() and (bid_p[:, 0] == 2_199_900).all()
    assert ((l2[:, 1::2] >= 1) & (l2[:, 1::2] <= 500)).all()
    assert np.array_equal(l2, synth.make_l2_books(32, seed=3))


def test_init_msgs_batched_matches_reference_recipe():
    l2 = synth.make_l2_books(5, seed=1)
    got = synth.init_msgs_from_l2_batched(l2, (34200, 7))
    for e in range(5):
        assert np.array_equal(got[e], O.init_msgs_from_l2(l2[e], (34200, 7)))


def test_stream_format_mixture_and_determinism():
    s1, s2 = synth.MessageStream(64, seed=5), synth.MessageStream(64, seed=5)
    m = s1.next(200)
    assert np.array_equal(m, s2.next(200)) and m.shape == (64, 200, 8) and m.dtype == np.int32
    t = m[..., 6].astype(np.int64) * 10 ** 9 + m[..., 7]
    assert (np.diff(t, axis=1) >= 0).all() and (m[..., 7] >= 0).all() and (m[..., 7] < 10 ** 9).all()
    assert set(np.unique(m[..., 0])) <= {1, 2, 3} and set(np.unique(m[..., 1])) <= {-1, 1}
    frac_limit = (m[..., 0] == 1).mean()
    assert 0.55 < frac_limit < 0.75
    lim = m[m[..., 0] == 1]
    assert (np.diff(np.sort(lim[:, 4])) >= 0).all() and (lim[:, 4] == lim[:, 5]).all()
    rows = synth.lobster_csv_rows(m[0, :3])
    assert len(rows) == 3 and len(rows[0]) == 6 and "." in rows[0][0]


def test_ffi_shim_source_is_guarded_and_names_the_abi():
    """csrc/ffi/vitmarl_ffi.cc (jax.ffi handlers) must compile to an empty translation unit where the XLA FFI headers are absent
    (this image) and bind exactly the C-ABI entry points the header declares."""
    import os, re, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "vitmarl_b200", "csrc", "ffi", "vitmarl_ffi.cc")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", src])
    text = open(src).read()
    header = open(os.path.join(root, "include", "vitmarl_b200.h")).read()
    for fn in set(re.findall(r"\b(vitmarl_[a-z0-9_]+)\(", text)):
        assert re.search(r"\b" + fn + r"\s*\(", header), fn
    from vitmarl_b200 import _build
    if _build.xla_ffi_include_dir() is None:
        import pytest
        with pytest.raises(RuntimeError):
            _build.build_ffi()
