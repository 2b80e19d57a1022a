"""The C-ABI library loads on a CPU-only host and exports every symbol include/octave_b200.h declares."""
import ctypes

from octave_b200 import _lib


def test_every_declared_symbol_is_exported():
    syms = _lib.declared_symbols()
    assert "octave_loss_fwd" in syms and len(syms) >= 6
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/octave_b200.h but not exported"


def test_abi_version_and_struct_size():
    assert _lib.lib.octave_abi_version() == 4
    # OctaveLossDesc: 7 int32 + 5 + 5 int32 + 4 float + 3 float + 2 int32 + 1 float (jsd_eps) = 27 words
    assert ctypes.sizeof(_lib.LossDesc) == 27 * 4


def test_argument_validation_without_gpu():
    d = _lib.LossDesc()
    d.dtype, d.B, d.C, d.H, d.W, d.flags = 7, 1, 2, 16, 16, _lib.LOSS_WPCE
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_INVALID
    d.dtype, d.C = 0, 99
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_INVALID
    d.C = 3  # generic path is fp32 only
    d.dtype = 1
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_UNSUPPORTED
    d.dtype, d.C = 0, 2
    assert _lib.lib.octave_loss_uses_fast_path(ctypes.byref(d)) == 1
    d.H = 20
    assert _lib.lib.octave_loss_uses_fast_path(ctypes.byref(d)) == 0
    assert _lib.lib.octave_loss_stats_bytes(ctypes.byref(d)) == (32 + 2) * 8


def test_new_entry_points_validate_arguments_without_gpu():
    # narrow-layer conv variants answer "unsupported" for descriptors outside their shape class, never crash
    from octave_b200._lib import ConvDesc
    d = ConvDesc()
    d.B, d.H, d.W, d.cin, d.cout, d.groups, d.ksize, d.stride, d.pad = 1, 8, 8, 48, 48, 1, 3, 1, 1
    d.Hout, d.Wout, d.x_ld, d.y_ld = 8, 8, 48, 48
    assert _lib.lib.octave_conv_halo_supported(ctypes.byref(d)) == 0          # 48 channels: not a {32, 64} layer
    assert _lib.lib.octave_conv_halo_wgrad_supported(ctypes.byref(d)) == 0
    d.cin, d.cout, d.x_ld, d.y_ld = 64, 32, 64, 32
    d.in_dtype = d.out_dtype = _lib.DTYPE_BF16
    assert _lib.lib.octave_conv_halo_supported(ctypes.byref(d)) == 1
    assert _lib.lib.octave_pack_job_blocks(0, 32, 64, 1, 3) == (32 * 64 + 1023) // 1024
    assert _lib.lib.octave_pack_weight_multi(None, 0, 0, None) == _lib.ERR_INVALID
