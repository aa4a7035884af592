"""ctypes binding of libvqb200.so (the C ABI in include/vqb200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is
raised.  The product path never routes through oracle/ or a CPU implementation.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# VQB_LIB_PATH: experiment builds of the same CUDA library (tools/ab_build.py); never a CPU fallback
LIB_PATH = os.environ.get("VQB_LIB_PATH") or os.path.join(_PKG, "libvqb200.so")

PATH_AUTO, PATH_FMA, PATH_TC = 0, 1, 2

_lib = None


class DeviceInfo(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("cc_major", ctypes.c_int), ("cc_minor", ctypes.c_int),
                ("sm_count", ctypes.c_int), ("max_smem_per_block", ctypes.c_int),
                ("l2_bytes", ctypes.c_size_t), ("total_mem", ctypes.c_size_t)]


# name -> (restype, argtypes); the single source for "every symbol include/vqb200.h declares"
_vp, _i, _i64, _f, _sz, _u = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float,
                              ctypes.c_size_t, ctypes.c_uint)
SIGNATURES = {
    "vqb_version": (_i, []),
    "vqb_error_string": (ctypes.c_char_p, [_i]),
    "vqb_query": (_i, [_i, ctypes.POINTER(DeviceInfo)]),
    "vqb_workspace_bytes": (_sz, [_i, _i]),
    "vqb_select_path": (_i, [_i, _i64, _i, _i, _i64, _i64]),
    "vqb_forward": (_i, [_i, _vp, _i64, _i64, _i, _i64, _i64, _i64, _vp, _i, _f,
                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp]),
    "vqb_backward": (_i, [_i, _vp, _vp, _vp, _i64, _i64, _i, _i64, _i64, _i64, _vp, _vp, _i, _f,
                          _vp, _vp, _vp, _sz, _vp]),
    "vqb_token_linear": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _u, _vp]),
    "vqb_token_linear_split": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _u, _i, _vp]),
    "vqb_token_conv_split": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _u, _i, _i, _i, _vp]),
    "vqb_token_out_proj_pair": (_i, [_i, _vp, _vp, _f, _vp, _i64, _i, _i, _i, _vp]),
    "vqb_token_pair": (_i, [_i, _vp, _vp, _i64, _i, _i, _vp]),
    "vqb_token_conv": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _u, _i, _i, _i, _vp]),
    "vqb_token_out_proj": (_i, [_i, _vp, _vp, _f, _vp, _i64, _i, _i, _vp]),
    "vqb_token_bias_gelu": (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp]),
    "vqb_encoder_chain_scratch_bytes": (_sz, [_i, _i]),
    "vqb_encoder_chain": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _sz, _vp, _vp, _i, _vp, _vp]),
    "vqb_patch_split": (_i, [_i, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "vqb_pack_rows": (_i, [_i, _vp, _i64, _i64, _i, _i64, _i64, _i64, _vp, _vp]),
    "vqb_patch_embed": (_i, [_i, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "vqb_ar_pairs": (_i, [_i, _vp, _i64, _i, _i64, _i64, _vp, _vp, _vp]),
    "vqb_row_keys": (_i, [_i, _vp, _i64, _i, _vp, _vp, _vp]),
    "vqb_dedupe_scratch_bytes": (_sz, [_i64]),
    "vqb_dedupe_first": (_i, [_i, _vp, _i64, _vp, _sz, _vp, _vp]),
    "vqb_gather": (_i, [_i, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp]),
    "vqb_one_hot": (_i, [_i, _vp, _i64, _i, _vp, _vp]),
    "vqb_launch_counter": (ctypes.c_longlong, []),
    "vqb_profile_enable": (_i, [_i]),
    "vqb_profile_collect": (_i, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i)]),
    "vqb_host_last_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "vqb_debug_set_tc_trace": (_i, [_vp]),
    "vqb_debug_tc_trace_words": (_sz, []),
    "vqb_debug_set_filter": (_i, [_i]),
    "vqb_host_create": (_i, [_i, _i64, _i, _i, _i, ctypes.POINTER(_vp)]),
    "vqb_host_destroy": (_i, [_vp]),
    "vqb_host_set_codebook": (_i, [_vp, _vp]),
    "vqb_encode_host": (_i, [_vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp, _u, ctypes.POINTER(_i)]),
}

E_UNSUPPORTED = -3      # VQB_E_UNSUPPORTED (include/vqb200.h)


def load():
    """Loads the CUDA library or raises.  No CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
            "(or vq-vae-transformer-arc-welding_b200/csrc/build.py). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def error_string(code: int) -> str:
    return load().vqb_error_string(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code}: {error_string(code)}")
