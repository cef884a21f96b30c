"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/vitmarl_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from vitmarl_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vitmarl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vitmarl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = _declared()
    assert "vitmarl_lob_step" in names and "vitmarl_env_step" in names
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_header_and_version():
    assert sorted(_capi.SIGNATURES) == _declared()
    assert _capi.lib().vitmarl_abi_version() == 3


def test_argument_validation_without_gpu():
    lib = _capi.lib()
    # unsupported cancel mode is rejected before any CUDA call (JaxOrderBookArrays.py:130-163 need jax.random)
    assert lib.vitmarl_lob_step(None, 1, 100, 100, 1, 1, None, None, None, None, None, None, None, None, None, 2, -2) == _capi.EUNSUPPORTED
    assert lib.vitmarl_lob_step(None, 1, 100, 100, 1, 1, None, None, None, None, None, None, None, None, None, 1, -2) == _capi.EINVAL
    assert lib.vitmarl_lob_step(None, 0, 100, 100, 1, 1, None, None, None, None, None, None, None, None, None, 1, -2) == _capi.OK
    assert lib.vitmarl_lob_render(None, 4, 300, 10, 100, None, None, None, None, None, None, None, 0, 0, 0) == _capi.EINVAL
    # struct-argument env step: same validation, reached through VitmarlEnvStepArgs
    a = _capi.EnvStepArgs()
    a.E, a.N, a.T, a.M, a.cancel_mode = 1, 100, 100, 13, 3
    assert lib.vitmarl_env_step2(None, ctypes.byref(a)) == _capi.EUNSUPPORTED
    a.cancel_mode = 1
    assert lib.vitmarl_env_step2(None, ctypes.byref(a)) == _capi.EINVAL         # null buffers
    assert lib.vitmarl_env_step2(None, None) == _capi.EINVAL
    # per-call options / timing handle instead of process-global switches
    t = lib.vitmarl_timing_create()
    assert t and lib.vitmarl_timing_reset(t) == _capi.OK and lib.vitmarl_timing_reset(None) == _capi.EINVAL
    lib.vitmarl_timing_destroy(t)
    shape = _capi.VitShape(4, 64, 64, 2, 8, 192, 12, 3, 768, 1e-6)
    assert lib.vitmarl_vit_num_buckets(ctypes.byref(shape)) == 14
    assert lib.vitmarl_vit_fwd_ex(None, ctypes.byref(shape), None, None, None, None, 0, 0, None) == _capi.EINVAL


def test_no_process_global_switches_in_the_abi():
    """the header promises no global mutable state on the compute path -- no set_* entry point may exist."""
    assert not [n for n in _declared() if "_set_" in n or n.endswith("_enable")]


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    from vitmarl_b200 import jaxob
    from vitmarl_b200.config import JAXLOB_Configuration
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = torch.full((1, 100, 6), -1, dtype=torch.int32)
    with pytest.raises(_capi.VitmarlError):
        jaxob.get_best_bid_and_ask_inclQuants(JAXLOB_Configuration(), a, a)
