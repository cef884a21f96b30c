"""ctypes binding of libvitmarl_b200.so (the C ABI declared in include/vitmarl_b200.h).

This is the reference-side binding a maintainer would write if JAX were absent; the
jax.ffi variant is shown in INTEGRATION.md.  No torch types cross the boundary: only raw
device pointers, sizes and a cudaStream_t."""
from __future__ import annotations

import ctypes
import os

from . import _build

OK, EINVAL, EUNSUPPORTED, ECUDA, ENODEVICE = 0, -1, -2, -3, -4
IMG_NONE, IMG_U8, IMG_BF16, IMG_BF16_PATCHES = 0, 1, 2, 3
VIT_INPUT_PATCHES = 4

_ERR_NAMES = {EINVAL: "VITMARL_EINVAL (bad shape / null / misaligned buffer)",
              EUNSUPPORTED: "VITMARL_EUNSUPPORTED (cancel_mode 2/3 or simulator_mode 1)",
              ECUDA: "VITMARL_ECUDA", ENODEVICE: "VITMARL_ENODEVICE"}


class VitmarlError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{_ERR_NAMES.get(code, code)} {detail}".strip())


class VitShape(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int), ("img_h", ctypes.c_int), ("img_w", ctypes.c_int),
                ("channels", ctypes.c_int), ("patch", ctypes.c_int), ("dim", ctypes.c_int),
                ("depth", ctypes.c_int), ("heads", ctypes.c_int), ("mlp_dim", ctypes.c_int),
                ("ln_eps", ctypes.c_float)]


class VitOptions(ctypes.Structure):
    """VitmarlVitOptions (include/vitmarl_b200.h); ``VitOptions.defaults()`` = every switch at its default."""
    _fields_ = [("fused", ctypes.c_int), ("gemm_2cta", ctypes.c_int), ("pdl", ctypes.c_int), ("attn_flags", ctypes.c_int),
                ("timing", ctypes.c_void_p), ("debug_timeline", ctypes.c_void_p),
                ("grads_flat", ctypes.c_void_p), ("grads_flat_bytes", ctypes.c_size_t), ("accumulate", ctypes.c_int),
                ("bucket_events", ctypes.c_void_p)]

    @classmethod
    def defaults(cls, **kw):
        o = cls(-1, -1, -1, -1, None, None, None, 0, -1, None)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class Timing:
    """Caller-owned CUDA-event log (vitmarl_timing_*): pass ``handle`` in ``VitOptions.timing``."""
    NAMES = ["gemm", "fused_mlp", "fused_attn_block", "attention", "layernorm", "other", "gemm_dW", "gemm_dX"]

    def __init__(self):
        self.handle = lib().vitmarl_timing_create()
        if not self.handle:
            raise VitmarlError(ECUDA, "vitmarl_timing_create failed")

    def reset(self):
        check(lib().vitmarl_timing_reset(self.handle))

    def read(self):
        ms, n, fl = (ctypes.c_double * 8)(), (ctypes.c_longlong * 8)(), ctypes.c_double()
        check(lib().vitmarl_timing_read(self.handle, ms, n, ctypes.byref(fl)))
        return list(ms), list(n), fl.value

    def close(self):
        if self.handle:
            lib().vitmarl_timing_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


GEMM_NO_2CTA = 1


class EnvStepArgs(ctypes.Structure):
    """VitmarlEnvStepArgs (include/vitmarl_b200.h)."""
    _fields_ = [("E", ctypes.c_int), ("N", ctypes.c_int), ("T", ctypes.c_int), ("M", ctypes.c_int),
                ("asks_in", ctypes.c_void_p), ("bids_in", ctypes.c_void_p), ("msgs", ctypes.c_void_p),
                ("last_ask_price", ctypes.c_void_p), ("last_bid_price", ctypes.c_void_p), ("last_price_stride", ctypes.c_int),
                ("asks_out", ctypes.c_void_p), ("bids_out", ctypes.c_void_p), ("trades_out", ctypes.c_void_p),
                ("best_asks", ctypes.c_void_p), ("best_bids", ctypes.c_void_p), ("mid_price", ctypes.c_void_p),
                ("n_levels", ctypes.c_int), ("tick_size", ctypes.c_int), ("raw", ctypes.c_void_p), ("l2", ctypes.c_void_p),
                ("norm", ctypes.c_void_p), ("image", ctypes.c_void_p), ("img_dtype", ctypes.c_int), ("H", ctypes.c_int),
                ("W", ctypes.c_int), ("cancel_mode", ctypes.c_int), ("init_id", ctypes.c_int32),
                ("n_stat_agents", ctypes.c_int), ("stat_agent_ids", ctypes.c_int32 * 4), ("trade_stats", ctypes.c_void_p),
                ("time_in", ctypes.c_void_p), ("time_out", ctypes.c_void_p), ("delta_time", ctypes.c_void_p)]


_lib = None
_P, _I, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/vitmarl_b200.h declares
SIGNATURES = {
    "vitmarl_abi_version": (_I, []),
    "vitmarl_last_error": (ctypes.c_char_p, []),
    "vitmarl_lob_step": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, ctypes.c_int32]),
    "vitmarl_lob_best_bid_ask": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "vitmarl_lob_render": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I]),
    "vitmarl_gemm_bf16": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, ctypes.c_float, _I]),
    "vitmarl_vit_num_params": (_I, [_P]),
    "vitmarl_vit_param_elems": (ctypes.c_longlong, [_P, _I, _P]),
    "vitmarl_vit_workspace_bytes": (_SZ, [_P, _I]),
    "vitmarl_vit_fwd": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _I]),
    "vitmarl_vit_bwd": (_I, [_P, _P, _P, _P, _SZ, _P, _P, _P]),
    "vitmarl_vit_fwd_ex": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _I, _P]),
    "vitmarl_vit_bwd_ex": (_I, [_P, _P, _P, _P, _SZ, _P, _P, _P, _P]),
    "vitmarl_vit_num_buckets": (_I, [_P]),
    "vitmarl_timing_create": (_P, []),
    "vitmarl_timing_destroy": (None, [_P]),
    "vitmarl_timing_reset": (_I, [_P]),
    "vitmarl_timing_read": (_I, [_P, _P, _P, _P]),
    "vitmarl_attention_fwd": (_I, [_P, _I, _I, _P, _P]),
    "vitmarl_attention_bwd": (_I, [_P, _I, _I, _P, _P, _P]),
    "vitmarl_dense_f32": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _I, _P, _I]),
    "vitmarl_gae_f32": (_I, [_P, _I, _I, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P, _P, _P]),
    "vitmarl_gru_cell_f32": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "vitmarl_debug_gemm_dw": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I]),
    "vitmarl_get_cancel_msgs": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "vitmarl_get_agent_trades": (_I, [_P, _I, _I, _P, _I, _P]),
    "vitmarl_agent_trade_stats": (_I, [_P, _I, _I, _P, _I, _I, _P]),
    "vitmarl_exec_action_msgs_fixed_quants_complex": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vitmarl_mm_action_msgs_spread_skew": (_I, [_P, _I, _P, _P, _P, _I, _P, _I, _I, ctypes.c_float, ctypes.c_float, _I, _I, _I, _I, _P]),
    "vitmarl_filter_messages": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "vitmarl_auto_reset": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vitmarl_env_step2": (_I, [_P, _P]),
    "vitmarl_build_step_msgs": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vitmarl_env_step": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                              _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, ctypes.c_int32]),
}


def library_path() -> str:
    return _build.LIB_PATH


def lib():
    """Load (building if the sources are newer) the CUDA library; raises if impossible."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if not os.path.exists(path) or os.environ.get("VITMARL_REBUILD"):
            path = _build.build()
        handle = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int):
    if code != OK:
        detail = lib().vitmarl_last_error().decode() if code == ECUDA else ""
        raise VitmarlError(code, detail)
