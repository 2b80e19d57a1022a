"""ctypes binding of liboctave_b200.so — the C-ABI in include/octave_b200.h.

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that
fails the import error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "lib" / "liboctave_b200.so"
_HEADER = _HERE.parent / "include" / "octave_b200.h"

OK = 0
ERR_INVALID = -1
ERR_UNSUPPORTED = -2
ERR_LAUNCH = -3

DTYPE_F32 = 0
DTYPE_BF16 = 1


class OctaveError(RuntimeError):
    """Non-zero return code of a C-ABI entry point."""

    def __init__(self, fn: str, rc: int):
        names = {ERR_INVALID: "OCT_ERR_INVALID", ERR_UNSUPPORTED: "OCT_ERR_UNSUPPORTED", ERR_LAUNCH: "OCT_ERR_LAUNCH"}
        super().__init__(f"{fn} failed with {names.get(rc, rc)}")
        self.rc = rc


def declared_symbols() -> list[str]:
    """Every function declared in include/octave_b200.h (used by the symbol-export test)."""
    text = _HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(octave_[a-z0-9_]+)\s*\(", text)))


def _load() -> C.CDLL:
    if not _LIB_PATH.exists() or os.environ.get("OCTAVE_B200_REBUILD"):
        from . import build as _build

        _build.build()
    return C.CDLL(str(_LIB_PATH))


lib = _load()


def check(fn: str, rc: int) -> None:
    if rc != OK:
        raise OctaveError(fn, rc)


# ---------------------------------------------------------------------------------------------
# struct mirrors
# ---------------------------------------------------------------------------------------------
class LossDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("flags", C.c_int32),
        ("n_att", C.c_int32),
        ("att_h", C.c_int32 * 5),
        ("att_w", C.c_int32 * 5),
        ("att_weight", C.c_float * 4),
        ("sum_weights", C.c_float),
        ("wpce_scale", C.c_float),
        ("dice_eps", C.c_float),
        ("n_real", C.c_int32), ("n_fake", C.c_int32),
        ("jsd_eps", C.c_float),
    ]


LOSS_WPCE = 1 << 0
LOSS_DICE = 1 << 1
LOSS_KLD = 1 << 2
LOSS_LSG = 1 << 3
LOSS_LSD = 1 << 4
LOSS_FROM_LOGITS = 1 << 5
LOSS_WPCE_FULL = 1 << 6
LOSS_KLD_STOPGRAD = 1 << 7
LOSS_JSD = 1 << 8
LOSS_OUT_TOTAL = 6
LOSS_OUT_SLOTS = 8

_vp = C.c_void_p

lib.octave_abi_version.restype = C.c_int
lib.octave_sm_count.restype = C.c_int
lib.octave_launch_count.restype = C.c_ulonglong
lib.octave_loss_stats_bytes.restype = C.c_size_t
lib.octave_loss_stats_bytes.argtypes = [C.POINTER(LossDesc)]
lib.octave_loss_uses_fast_path.restype = C.c_int
lib.octave_loss_uses_fast_path.argtypes = [C.POINTER(LossDesc)]
lib.octave_loss_fwd.restype = C.c_int
lib.octave_loss_fwd.argtypes = [C.POINTER(LossDesc), _vp, _vp, C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp]
lib.octave_loss_bwd.restype = C.c_int
lib.octave_loss_bwd.argtypes = [C.POINTER(LossDesc), _vp, _vp, C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp,
                                C.POINTER(_vp), _vp, _vp, _vp]
lib.octave_loss_fused_supported.restype = C.c_int
lib.octave_loss_fused_supported.argtypes = [C.POINTER(LossDesc)]
lib.octave_loss_fused_stats_bytes.restype = C.c_size_t
lib.octave_loss_fused_stats_bytes.argtypes = [C.POINTER(LossDesc)]
lib.octave_loss_fused.restype = C.c_int
lib.octave_loss_fused.argtypes = [C.POINTER(LossDesc), _vp, _vp, C.POINTER(_vp), _vp, C.POINTER(C.c_float), _vp, _vp, _vp,
                                  C.POINTER(_vp), _vp, _vp]
lib.octave_wpce_alt_fwd.restype = C.c_int
lib.octave_wpce_alt_fwd.argtypes = [C.c_int32, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]
lib.octave_wpce_alt_bwd.restype = C.c_int
lib.octave_wpce_alt_bwd.argtypes = [C.c_int32, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]
lib.octave_loss_scale_grads.restype = C.c_int
lib.octave_loss_scale_grads.argtypes = [C.POINTER(LossDesc), _vp, _vp, C.POINTER(_vp), _vp, _vp]


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32), ("groups", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("x_ld", C.c_int32), ("x_coff", C.c_int32),
        ("y_ld", C.c_int32), ("y_coff", C.c_int32),
        ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("mode", C.c_int32), ("relu", C.c_int32),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("accumulate", C.c_int32), ("real_groups", C.c_int32), ("out_s2d_qs", C.c_int32),
    ]


CONV_MODE_CONV = 0
CONV_MODE_CONVT = 1

lib.octave_conv_tc_supported.restype = C.c_int
lib.octave_conv_tc_supported.argtypes = [C.POINTER(ConvDesc)]
lib.octave_conv_tc_fwd.restype = C.c_int
lib.octave_conv_tc_fwd.argtypes = [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp]
lib.octave_conv_tc_wgrad_supported.restype = C.c_int
lib.octave_conv_tc_wgrad_supported.argtypes = [C.POINTER(ConvDesc)]
lib.octave_conv_tc_wgrad.restype = C.c_int
lib.octave_conv_tc_wgrad.argtypes = [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]


# ---- bf16 gradient buckets (include/octave_b200.h: octave_grad_pack_bf16 / octave_grad_unpack_bf16) ----
class OctaveGradJob(C.Structure):
    _fields_ = [("g", C.c_void_p), ("flat_off", C.c_int64), ("n", C.c_int32), ("block_start", C.c_int32)]


GRAD_MAX_JOBS = 128
lib.octave_optim_job_blocks.restype = C.c_int64
lib.octave_optim_job_blocks.argtypes = [C.c_int64]
for _n in ("octave_grad_pack_bf16", "octave_grad_unpack_bf16"):
    getattr(lib, _n).restype = C.c_int
    getattr(lib, _n).argtypes = [C.POINTER(OctaveGradJob), C.c_int32, C.c_void_p, C.c_void_p]
