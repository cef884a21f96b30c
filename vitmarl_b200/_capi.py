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


_lib = None
_P, _I, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/vitmarl_b200.h declares
SIGNATURES = {
    "vitmarl_abi_version": (_I, []),
    "vitmarl_last_error": (ctypes.c_char_p, []),
    "vitmarl_lob_step": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, ctypes.c_int32]),
    "vitmarl_lob_best_bid_ask": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "vitmarl_lob_render": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I]),
    "vitmarl_gemm_bf16": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, ctypes.c_float]),
    "vitmarl_gemm_set_2cta": (_I, [_I]),
    "vitmarl_vit_num_params": (_I, [_P]),
    "vitmarl_vit_param_elems": (ctypes.c_longlong, [_P, _I, _P]),
    "vitmarl_vit_workspace_bytes": (_SZ, [_P, _I]),
    "vitmarl_vit_fwd": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _I]),
    "vitmarl_vit_bwd": (_I, [_P, _P, _P, _P, _SZ, _P, _P, _P]),
    "vitmarl_vit_set_fused": (_I, [_I]),
    "vitmarl_debug_fused_mlp_timeline": (_I, [_P]),
    "vitmarl_debug_set_flags": (_I, [_I]),
    "vitmarl_debug_gemm_dw": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "vitmarl_vit_gemm_timing_enable": (_I, [_I]),
    "vitmarl_vit_gemm_timing_read": (_I, [_P, _P, _P]),
    "vitmarl_vit_timing_read_categories": (_I, [_P, _P]),
    "vitmarl_get_cancel_msgs": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "vitmarl_get_agent_trades": (_I, [_P, _I, _I, _P, _I, _P]),
    "vitmarl_agent_trade_stats": (_I, [_P, _I, _I, _P, _I, _I, _P]),
    "vitmarl_filter_messages": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "vitmarl_auto_reset": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
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
