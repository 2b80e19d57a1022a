"""Thin Python wrappers over the C-ABI (include/octave_b200.h).  Everything here works on `Act`, an NHWC
channel view of a torch buffer; torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import lib

_vp = C.c_void_p


def stream_ptr() -> int:
    # raw cudaStream_t of torch's current stream (the fast C accessor; torch.cuda.current_stream() costs ~14 us)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _dt(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return _lib.DTYPE_F32
    if dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise TypeError(f"octave_b200 supports float32 / bfloat16 activations, got {dtype}")


class ActStruct(C.Structure):
    _fields_ = [("data", _vp), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("ld", C.c_int32), ("coff", C.c_int32), ("dtype", C.c_int32)]


class PoolDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("ceil_mode", C.c_int32), ("count_include_pad", C.c_int32)]


class PackJob(C.Structure):
    _fields_ = [("w", _vp), ("out", _vp), ("mode", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32), ("groups", C.c_int32),
                ("dense_groups", C.c_int32), ("ksize", C.c_int32), ("block_start", C.c_int64)]


class Act:
    """NHWC channel view: pixel p, channel c at buf.data_ptr + (p*ld + coff + c) elements."""
    __slots__ = ("buf", "B", "H", "W", "C", "ld", "coff")

    def __init__(self, buf: torch.Tensor, B: int, H: int, W: int, C_: int, ld: Optional[int] = None, coff: int = 0):
        self.buf, self.B, self.H, self.W, self.C = buf, B, H, W, C_
        self.ld = C_ if ld is None else ld
        self.coff = coff

    @staticmethod
    def empty(B: int, H: int, W: int, C_: int, dtype: torch.dtype, device) -> "Act":
        return Act(torch.empty((B, H, W, C_), dtype=dtype, device=device), B, H, W, C_)

    @staticmethod
    def zeros(B: int, H: int, W: int, C_: int, dtype: torch.dtype, device) -> "Act":
        return Act(torch.zeros((B, H, W, C_), dtype=dtype, device=device), B, H, W, C_)

    def like(self, C_: Optional[int] = None) -> "Act":
        return Act.empty(self.B, self.H, self.W, self.C if C_ is None else C_, self.buf.dtype, self.buf.device)

    def slice(self, c0: int, c: int) -> "Act":
        assert c0 + c <= self.C
        return Act(self.buf, self.B, self.H, self.W, c, self.ld, self.coff + c0)

    @property
    def dtype(self) -> torch.dtype:
        return self.buf.dtype

    @property
    def device(self):
        return self.buf.device

    @property
    def npix(self) -> int:
        return self.B * self.H * self.W

    def struct(self) -> ActStruct:
        return ActStruct(self.buf.data_ptr(), self.B, self.H, self.W, self.C, self.ld, self.coff, _dt(self.buf.dtype))

    def to_nchw(self) -> torch.Tensor:
        """Debug/test helper (torch ops): materialise the view as an NCHW fp32 tensor."""
        t = self.buf.reshape(self.B, self.H, self.W, self.ld)[..., self.coff:self.coff + self.C]
        return t.permute(0, 3, 1, 2).float().contiguous()


def _ref(a: Optional[Act]):
    return None if a is None else C.byref(a.struct())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(name: str, rc: int) -> None:
    if rc != 0:
        raise _lib.OctaveError(name, rc)


# --- bind the remaining entry points ---------------------------------------------------------------
_A = C.POINTER(ActStruct)
_sigs = {
    "octave_conv_direct_fwd": [C.POINTER(_lib.ConvDesc), _vp, _vp, _vp, _vp, _vp],
    "octave_conv_direct_dgrad": [C.POINTER(_lib.ConvDesc), _vp, _vp, _vp, _vp],
    "octave_conv_direct_wgrad": [C.POINTER(_lib.ConvDesc), _vp, _vp, _vp, _vp, _vp],
    "octave_act_bwd": [_A, _A, C.c_int32, _A, _vp],
    "octave_chan_stats": [_A, _vp, _vp],
    "octave_bn_prepare": [C.c_int32, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_int32, _vp, _vp, _vp],
    "octave_affine_act": [_A, _vp, _A, C.c_int32, _A, _vp, _vp, _vp],
    "octave_bn_bwd_reduce": [_A, _A, _vp, _A, _vp, _vp, _A, _vp],
    "octave_bn_bwd_apply": [_A, _A, _vp, _A, _vp, _vp, _vp, C.c_int32, _A, _vp, _vp, _A, _vp],
    "octave_add_inplace": [_A, _A, _vp],
    "octave_relu_bwd": [_A, _A, _A, _vp],
    "octave_splat_combine": [_A, _vp, C.c_int32, _A, _vp],
    "octave_splat_bwd_reduce": [_A, _A, _A, _vp, _vp],
    "octave_splat_bwd_du": [_A, _A, _vp, _vp, C.c_float, _A, _vp],
    "octave_splat_bn_bwd": [_A, _A, _vp, _vp, C.c_float, _A, _vp, _vp, _vp, C.c_int32, _vp, _A, _vp, _vp, _vp],
    "octave_pool_out_size": [C.POINTER(PoolDesc), C.c_int32],
    "octave_pool_fwd": [C.POINTER(PoolDesc), _A, _A, _vp, _vp],
    "octave_pool_bwd": [C.POINTER(PoolDesc), _A, _vp, _A, _vp],
    "octave_head_fwd": [_A, _vp, _vp, C.c_int32, C.c_int32, _vp, _A, _vp],
    "octave_head_bwd": [_A, _vp, _vp, C.c_int32, C.c_int32, _vp, _A, _A, _vp, _vp, _vp, _vp],
    "octave_head_wgrad": [_A, _vp, C.c_int32, _vp, _vp, _vp],
    "octave_nchw_to_nhwc": [_vp, C.c_int32, _A, _vp],
    "octave_nhwc_to_nchw": [_A, _vp, C.c_int32, _vp],
    "octave_nchw_to_nhwc_noise": [_vp, C.c_int32, _vp, C.c_int32, _A, _vp],
    "octave_nhwc_to_nchw_clipmask": [_A, _vp, _vp, C.c_int32, _vp, _vp],
    "octave_copy_window": [_A, _A, C.c_int32, _vp],
    "octave_space_to_depth": [_A, _A, _vp, _vp],
    "octave_depth_to_space": [_A, _A, _vp],
    "octave_nchw_to_s2d": [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int32, _A, C.c_int32, C.c_int32, _vp],
    "octave_s2d_to_nchw": [_A, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, _vp, _vp],
    "octave_pack_weight_s2d": [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_unpack_wgrad_s2d": [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_rowdot_fwd": [_A, _vp, _vp, _vp, _vp],
    "octave_rowdot_bwd": [_A, _vp, _vp, _A, _vp, _vp, _vp],
    "octave_glinear_fwd": [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _vp, _vp],
    "octave_glinear_bwd_data": [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _vp, _vp],
    "octave_glinear_bwd_weight": [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _vp, _vp, _vp],
    "octave_bn1d_relu_fwd": [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_int32, _vp, _vp, _vp],
    "octave_bn1d_relu_bwd": [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp],
    "octave_rsoftmax_fwd": [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_rsoftmax_bwd": [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_attn_fused_supported": [C.c_int32] * 5,
    "octave_glinear_bn_relu_fwd": [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float,
                                   C.c_int32, _vp, _vp, _vp, _vp],
    "octave_glinear_rsoftmax_fwd": [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_rsoftmax_glinear_bn_bwd": [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp],
    "octave_pack_weight": [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp],
    "octave_pack_weight_multi": [_vp, C.c_int32, C.c_int64, _vp],
}
for _n, _a in _sigs.items():
    getattr(lib, _n).restype = C.c_int
    getattr(lib, _n).argtypes = _a
lib.octave_spectral_sigma.restype = C.c_int
lib.octave_spectral_sigma.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, C.c_float, _vp, _vp]
class OctaveSnJob(C.Structure):
    _fields_ = [("W", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32)]


SN_MAX_JOBS = 8
lib.octave_spectral_sigma_multi.restype = C.c_int
lib.octave_spectral_sigma_multi.argtypes = [C.POINTER(OctaveSnJob), C.c_int32, C.c_int32, C.c_float, _vp]
lib.octave_spectral_wgrad.restype = C.c_int
lib.octave_spectral_wgrad.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, _vp]
lib.octave_affine_gap_ws_bytes.restype = C.c_size_t
lib.octave_affine_gap_ws_bytes.argtypes = [_A]
lib.octave_pack_job_blocks.restype = C.c_int64
lib.octave_pack_job_blocks.argtypes = [C.c_int32] * 5

lib.octave_set_stats_prezeroed.restype = None
lib.octave_set_stats_prezeroed.argtypes = [C.c_int]
lib.octave_stream_capture_id.restype = C.c_ulonglong
lib.octave_stream_capture_id.argtypes = [_vp]

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4

# --- pre-zeroed statistics arena --------------------------------------------------------------------
# The double-precision statistics outputs (conv forward `stats`, BatchNorm `sums` / `sums2`, space-to-depth `chan_sum`)
# are accumulated with atomics, so they must start at zero: ~250 per training step, each a memset node of a few KB in
# front of its kernel.  They are slices of ONE chunk zeroed by one memset instead (include/octave_b200.h:
# octave_set_stats_prezeroed).  A chunk belongs to (device, stream, capture): a slice is only ever used on the stream that
# zeroed the chunk, and a chunk zeroed outside a graph capture is never handed out inside it (a replay would find the
# sums of the previous replay), nor one capture's chunk in the next.  Exhausted chunks stay alive through their slices.
_ZCHUNK = 1 << 16                  # doubles per chunk (512 KB)
_zero_pool = {}                    # (device index, stream handle) -> [capture id, chunk, next free offset]
_PREZERO = os.environ.get("OCTAVE_STATS_ARENA", "1") != "0"      # 0: per-call memsets (A/B switch)
lib.octave_set_stats_prezeroed(1 if _PREZERO else 0)


def zeros_f64(n: int, device) -> torch.Tensor:
    if not _PREZERO:
        return torch.empty(n, dtype=torch.float64, device=device)
    sp = stream_ptr()
    cap = int(lib.octave_stream_capture_id(sp))
    key = (device.index, sp)
    ent = _zero_pool.get(key)
    n_al = (n + 15) & ~15          # slices start on 128-byte lines
    if ent is None or ent[0] != cap or ent[2] + n_al > ent[1].numel():
        ent = [cap, torch.zeros(max(_ZCHUNK, n_al), dtype=torch.float64, device=device), 0]
        _zero_pool[key] = ent
    out = ent[1][ent[2]:ent[2] + n]
    ent[2] += n_al
    return out



# --- BatchNorm -------------------------------------------------------------------------------------
def chan_stats(x: Act) -> torch.Tensor:
    sums = zeros_f64(2 * x.C, x.device)
    _chk("octave_chan_stats", lib.octave_chan_stats(_ref(x), sums.data_ptr(), stream_ptr()))
    return sums


def bn_prepare(C_: int, count: float, sums, gamma, beta, rm, rv, nbt, eps: float, momentum: float, training: bool, device):
    ab = torch.empty(2 * C_, dtype=torch.float32, device=device)
    mi = torch.empty(2 * C_, dtype=torch.float32, device=device)
    _chk("octave_bn_prepare", lib.octave_bn_prepare(C_, float(count), _p(sums), _p(gamma), _p(beta), _p(rm), _p(rv), _p(nbt),
                                                    eps, momentum, int(training), ab.data_ptr(), mi.data_ptr(), stream_ptr()))
    return ab, mi


def affine_act(x: Act, ab: Optional[torch.Tensor], res: Optional[Act], relu: bool, out: Optional[Act] = None,
               want_gap: bool = False):
    y = out if out is not None else x.like()
    gap = ws = None
    if want_gap:
        gap = torch.empty((x.B, x.C // 2), dtype=torch.float32, device=x.device)
        ws = _gap_workspace(x)
    _chk("octave_affine_act", lib.octave_affine_act(_ref(x), _p(ab), _ref(res), int(relu), _ref(y), _p(gap), _p(ws), stream_ptr()))
    return y, gap


_gap_ws = {}


def _gap_workspace(x: Act) -> torch.Tensor:
    """Scratch of the fixed-order global-average-pool reduction; one per (device, stream), zeroed once (its counters are
    left at zero by every launch, and launches on one stream are serialised)."""
    nbytes = lib.octave_affine_gap_ws_bytes(_ref(x))
    key = (x.device, stream_ptr())
    ws = _gap_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 16 << 20), dtype=torch.uint8, device=x.device)
        _gap_ws[key] = ws
    return ws


def bn_bwd(dy: Act, mask: Optional[Act], x: Act, mi: torch.Tensor, gamma: Optional[torch.Tensor], training: bool,
           out: Optional[Act] = None, relu_ab: Optional[torch.Tensor] = None, dmasked: Optional[Act] = None):
    """-> dx, dgamma, dbeta.  relu_ab (with mask None): recompute this BN's own ReLU mask from x instead of reading it.
    dmasked: also receives dy * (mask > 0) (the residual branch's gradient)."""
    sums2 = zeros_f64(2 * x.C, x.device)
    # dmasked leaves with the reduction pass; the apply pass then reads it in place of dy AND the mask
    _chk("octave_bn_bwd_reduce", lib.octave_bn_bwd_reduce(_ref(dy), _ref(mask), _p(relu_ab), _ref(x), mi.data_ptr(), sums2.data_ptr(),
                                                          _ref(dmasked), stream_ptr()))
    dx = out if out is not None else x.like()
    dgamma = torch.empty(x.C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(x.C, dtype=torch.float32, device=x.device)
    if dmasked is not None:
        dy, mask, relu_ab = dmasked, None, None
    _chk("octave_bn_bwd_apply", lib.octave_bn_bwd_apply(_ref(dy), _ref(mask), _p(relu_ab), _ref(x), mi.data_ptr(), _p(gamma), sums2.data_ptr(),
                                                        int(training), _ref(dx), dgamma.data_ptr(), dbeta.data_ptr(), None, stream_ptr()))
    return dx, dgamma, dbeta


def add_inplace(dst: Act, src: Act) -> None:
    _chk("octave_add_inplace", lib.octave_add_inplace(_ref(dst), _ref(src), stream_ptr()))


def relu_bwd(dy: Act, mask: Act, out: Optional[Act] = None) -> Act:
    dx = out if out is not None else dy.like()
    _chk("octave_relu_bwd", lib.octave_relu_bwd(_ref(dy), _ref(mask), _ref(dx), stream_ptr()))
    return dx


def act_bwd(y: Act, dy: Act, act: int, out: Optional[Act] = None) -> Act:
    dz = out if out is not None else dy.like()
    _chk("octave_act_bwd", lib.octave_act_bwd(_ref(y), _ref(dy), act, _ref(dz), stream_ptr()))
    return dz


# --- split attention ---------------------------------------------------------------------------------
def splat_combine(U: Act, att: torch.Tensor, relu: bool, out: Optional[Act] = None) -> Act:
    o = out if out is not None else U.like(U.C // 2)
    _chk("octave_splat_combine", lib.octave_splat_combine(_ref(U), att.data_ptr(), int(relu), _ref(o), stream_ptr()))
    return o


def splat_bwd_reduce(dout: Act, mask: Optional[Act], U: Act) -> torch.Tensor:
    datt = torch.empty((dout.B, 2 * dout.C), dtype=torch.float32, device=dout.device)
    _chk("octave_splat_bwd_reduce", lib.octave_splat_bwd_reduce(_ref(dout), _ref(mask), _ref(U), datt.data_ptr(), stream_ptr()))
    return datt


def splat_bwd_du(dout: Act, mask: Optional[Act], att: torch.Tensor, dgap: Optional[torch.Tensor], gap_scale: float) -> Act:
    dU = dout.like(2 * dout.C)
    _chk("octave_splat_bwd_du", lib.octave_splat_bwd_du(_ref(dout), _ref(mask), att.data_ptr(), _p(dgap), gap_scale, _ref(dU), stream_ptr()))
    return dU


def splat_bn_bwd(dout: Act, omask: Optional[Act], att: torch.Tensor, dgap: Optional[torch.Tensor], gap_scale: float, z: Act,
                 ab: torch.Tensor, mi: torch.Tensor, gamma: torch.Tensor, training: bool):
    """Fused backward of the split-attention combine + bn0 + ReLU -> dz [B,H,W,2C], dgamma, dbeta (see the C header)."""
    dz = z.like()
    sums2 = zeros_f64(2 * z.C, z.device)
    dg = torch.empty(z.C, dtype=torch.float32, device=z.device)
    db = torch.empty(z.C, dtype=torch.float32, device=z.device)
    _chk("octave_splat_bn_bwd", lib.octave_splat_bn_bwd(_ref(dout), _ref(omask), att.data_ptr(), _p(dgap), gap_scale, _ref(z), ab.data_ptr(),
                                                        mi.data_ptr(), _p(gamma), int(training), sums2.data_ptr(), _ref(dz), dg.data_ptr(),
                                                        db.data_ptr(), stream_ptr()))
    return dz, dg, db


def glinear_fwd(inp: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], groups: int, in_scale: float) -> torch.Tensor:
    B, Kt = inp.shape
    N = w.shape[0]
    out = torch.empty((B, N), dtype=torch.float32, device=inp.device)
    _chk("octave_glinear_fwd", lib.octave_glinear_fwd(inp.data_ptr(), w.data_ptr(), _p(bias), B, Kt, N, groups, in_scale, out.data_ptr(), stream_ptr()))
    return out


def glinear_bwd(dout: torch.Tensor, inp: torch.Tensor, w: torch.Tensor, groups: int, in_scale: float, side=None):
    """-> din, dw (same shape as w), dbias.  side: optional `run(fn, *tensors)` that launches the parameter-gradient kernel on
    another stream (network.on_side_stream): only din is on the critical path of the backward pass."""
    B, Kt = inp.shape
    N = w.shape[0]
    din = torch.empty_like(inp)
    _chk("octave_glinear_bwd_data", lib.octave_glinear_bwd_data(dout.data_ptr(), w.data_ptr(), B, Kt, N, groups, in_scale, din.data_ptr(), stream_ptr()))

    def weight_part():
        dw_ = torch.empty_like(w)
        db_ = torch.empty(N, dtype=torch.float32, device=inp.device)
        _chk("octave_glinear_bwd_weight", lib.octave_glinear_bwd_weight(dout.data_ptr(), inp.data_ptr(), B, Kt, N, groups, in_scale, dw_.data_ptr(), db_.data_ptr(), stream_ptr()))
        return dw_, db_

    dw, db = side(weight_part, dout, inp) if side is not None else weight_part()
    return din, dw, db


def attn_fused_ok(B: int, C_: int, inter: int, groups: int, radix: int) -> bool:
    from . import config
    return bool(config.fuse_attention_branch and lib.octave_attn_fused_supported(B, C_, inter, groups, radix))


def glinear_bn_relu_fwd(inp, w, bias, in_scale: float, gamma, beta, rm, rv, nbt, eps: float, momentum: float, training: bool):
    """fc1 + BatchNorm1d + ReLU in one launch -> x (pre-BN, kept for backward), y, mean_invstd."""
    B, Kt = inp.shape
    N = w.shape[0]
    x = torch.empty((B, N), dtype=torch.float32, device=inp.device)
    y = torch.empty_like(x)
    mi = torch.empty(2 * N, dtype=torch.float32, device=inp.device)
    _chk("octave_glinear_bn_relu_fwd", lib.octave_glinear_bn_relu_fwd(
        inp.data_ptr(), w.data_ptr(), _p(bias), B, Kt, N, in_scale, gamma.data_ptr(), beta.data_ptr(), _p(rm), _p(rv), _p(nbt), eps, momentum,
        int(training), x.data_ptr(), y.data_ptr(), mi.data_ptr(), stream_ptr()))
    return x, y, mi


def glinear_rsoftmax_fwd(inp, w, bias, C_: int) -> torch.Tensor:
    """fc2 + r-softmax over the radix pair (c, C + c) in one launch -> att [B, 2C]."""
    B, Kt = inp.shape
    att = torch.empty((B, 2 * C_), dtype=torch.float32, device=inp.device)
    _chk("octave_glinear_rsoftmax_fwd", lib.octave_glinear_rsoftmax_fwd(inp.data_ptr(), w.data_ptr(), _p(bias), B, Kt, C_, att.data_ptr(), stream_ptr()))
    return att


def attn_bwd_fused(datt, att, w2, h1, h1n, gamma, mi, training: bool, groups: int, side=None):
    """r-softmax backward + fc2 data gradient + BatchNorm1d/ReLU backward in one launch; fc2's parameter gradients from the
    dlogits it leaves behind (on `side`, see glinear_bwd).  -> dx (gradient of fc1's output), dgamma, dbeta, dw2, db2"""
    B, Kt = h1.shape
    C_ = att.shape[1] // 2
    dlogits = torch.empty_like(att)
    dx = torch.empty_like(h1)
    dg = torch.empty(Kt, dtype=torch.float32, device=h1.device)
    db = torch.empty(Kt, dtype=torch.float32, device=h1.device)
    _chk("octave_rsoftmax_glinear_bn_bwd", lib.octave_rsoftmax_glinear_bn_bwd(
        datt.data_ptr(), att.data_ptr(), w2.data_ptr(), B, Kt, C_, h1.data_ptr(), h1n.data_ptr(), gamma.data_ptr(), mi.data_ptr(), int(training),
        dlogits.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), stream_ptr()))

    def weight_part():
        dw_ = torch.empty_like(w2)
        db_ = torch.empty(2 * C_, dtype=torch.float32, device=h1.device)
        _chk("octave_glinear_bwd_weight", lib.octave_glinear_bwd_weight(dlogits.data_ptr(), h1n.data_ptr(), B, Kt, 2 * C_, groups, 1.0, dw_.data_ptr(), db_.data_ptr(), stream_ptr()))
        return dw_, db_

    dw2, db2 = side(weight_part, dlogits, h1n) if side is not None else weight_part()
    return dx, dg, db, dw2, db2


def glinear_bwd_data_only(dout: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """din[b][i] = sum_j dout[b][j] * w[j][i]  (W^T applied to rows of dout)"""
    B, N = dout.shape
    Kt = w.shape[1]
    din = torch.empty((B, Kt), dtype=torch.float32, device=w.device)
    d = dout.contiguous().float()
    _chk("octave_glinear_bwd_data", lib.octave_glinear_bwd_data(d.data_ptr(), w.data_ptr(), B, Kt, N, 1, 1.0, din.data_ptr(), stream_ptr()))
    return din


def bn1d_relu_fwd(x: torch.Tensor, gamma, beta, rm, rv, nbt, eps: float, momentum: float, training: bool):
    B, C_ = x.shape
    y = torch.empty_like(x)
    mi = torch.empty(2 * C_, dtype=torch.float32, device=x.device)
    _chk("octave_bn1d_relu_fwd", lib.octave_bn1d_relu_fwd(x.data_ptr(), B, C_, gamma.data_ptr(), beta.data_ptr(), _p(rm), _p(rv), _p(nbt),
                                                          eps, momentum, int(training), y.data_ptr(), mi.data_ptr(), stream_ptr()))
    return y, mi


def bn1d_relu_bwd(dy, x, y, gamma, mi, training: bool):
    B, C_ = x.shape
    dx = torch.empty_like(x)
    dg = torch.empty(C_, dtype=torch.float32, device=x.device)
    db = torch.empty(C_, dtype=torch.float32, device=x.device)
    _chk("octave_bn1d_relu_bwd", lib.octave_bn1d_relu_bwd(dy.data_ptr(), x.data_ptr(), y.data_ptr(), B, C_, gamma.data_ptr(), mi.data_ptr(),
                                                          int(training), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), stream_ptr()))
    return dx, dg, db


def rsoftmax_fwd(logits: torch.Tensor, R: int) -> torch.Tensor:
    B, RC = logits.shape
    att = torch.empty_like(logits)
    _chk("octave_rsoftmax_fwd", lib.octave_rsoftmax_fwd(logits.data_ptr(), B, R, RC // R, att.data_ptr(), stream_ptr()))
    return att


def rsoftmax_bwd(datt: torch.Tensor, att: torch.Tensor, R: int) -> torch.Tensor:
    B, RC = att.shape
    dl = torch.empty_like(att)
    _chk("octave_rsoftmax_bwd", lib.octave_rsoftmax_bwd(datt.data_ptr(), att.data_ptr(), B, R, RC // R, dl.data_ptr(), stream_ptr()))
    return dl


# --- pools ---------------------------------------------------------------------------------------------
def pool_desc(kind: str, k: int, stride: int, pad: int, ceil_mode: bool = False, count_include_pad: bool = True) -> PoolDesc:
    return PoolDesc(0 if kind == "max" else 1, k, stride, pad, int(ceil_mode), int(count_include_pad))


def pool_fwd(pd: PoolDesc, x: Act):
    Ho = lib.octave_pool_out_size(C.byref(pd), x.H)
    Wo = lib.octave_pool_out_size(C.byref(pd), x.W)
    y = Act.empty(x.B, Ho, Wo, x.C, x.dtype, x.device)
    arg = torch.empty((x.B, Ho, Wo, x.C), dtype=torch.uint8, device=x.device) if pd.kind == 0 else None
    _chk("octave_pool_fwd", lib.octave_pool_fwd(C.byref(pd), _ref(x), _ref(y), _p(arg), stream_ptr()))
    return y, arg


def pool_bwd(pd: PoolDesc, dy: Act, arg: Optional[torch.Tensor], H: int, W: int) -> Act:
    dx = Act.empty(dy.B, H, W, dy.C, dy.dtype, dy.device)
    _chk("octave_pool_bwd", lib.octave_pool_bwd(C.byref(pd), _ref(dy), _p(arg), _ref(dx), stream_ptr()))
    return dx


# --- heads ---------------------------------------------------------------------------------------------
def head_fwd(x: Act, w: torch.Tensor, b: torch.Tensor, mode: int):
    """w [K, C] fp32.  -> (out [B,K,H,W] fp32, gated Act or None)"""
    K = w.shape[0]
    out = torch.empty((x.B, K, x.H, x.W), dtype=torch.float32, device=x.device)
    gated = x.like() if mode == 1 else None
    _chk("octave_head_fwd", lib.octave_head_fwd(_ref(x), w.data_ptr(), b.data_ptr(), K, mode, out.data_ptr(), _ref(gated), stream_ptr()))
    return out, gated


def head_bwd(x: Act, w: torch.Tensor, b: torch.Tensor, mode: int, dout: Optional[torch.Tensor], dgated: Optional[Act]):
    """-> dx Act, dw [K,C], db [K]"""
    K = w.shape[0]
    dx = x.like()
    dw = torch.empty((K, x.C), dtype=torch.float32, device=x.device)
    db = torch.empty(K, dtype=torch.float32, device=x.device)
    # K == 2: one fused pass (dx + parameter gradients); otherwise the generic kernel pair through a dlogits scratch
    dlogits = None if K == 2 and x.C >= 32 else torch.empty((x.B, K, x.H, x.W), dtype=torch.float32, device=x.device)
    _chk("octave_head_bwd", lib.octave_head_bwd(_ref(x), w.data_ptr(), b.data_ptr(), K, mode, _p(dout), _ref(dgated), _ref(dx),
                                                _p(dlogits), dw.data_ptr(), db.data_ptr(), stream_ptr()))
    return dx, dw, db


# --- layout ----------------------------------------------------------------------------------------------
def nchw_to_nhwc(src: torch.Tensor, dtype: torch.dtype, pad_to: int = 8) -> Act:
    B, Cs, H, W = src.shape
    Cd = ((Cs + pad_to - 1) // pad_to) * pad_to
    dst = Act.empty(B, H, W, Cd, dtype, src.device)
    s = src.contiguous().float()
    _chk("octave_nchw_to_nhwc", lib.octave_nchw_to_nhwc(s.data_ptr(), Cs, _ref(dst), stream_ptr()))
    return Act(dst.buf, B, H, W, Cs, Cd, 0)


def nhwc_to_nchw(src: Act, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    dst = out if out is not None else torch.empty((src.B, src.C, src.H, src.W), dtype=torch.float32, device=src.device)
    _chk("octave_nhwc_to_nchw", lib.octave_nhwc_to_nchw(_ref(src), dst.data_ptr(), int(accumulate), stream_ptr()))
    return dst


def nchw_into(src: torch.Tensor, dst: Act, noise: Optional[torch.Tensor] = None, clip: bool = False) -> None:
    """Write an NCHW fp32 tensor into the channel view `dst` (dst.C >= src channels), optionally + noise plane and clip."""
    s = src.contiguous().float()
    _chk("octave_nchw_to_nhwc_noise", lib.octave_nchw_to_nhwc_noise(s.data_ptr(), s.shape[1], _p(noise), int(clip), _ref(dst), stream_ptr()))


def nhwc_to_nchw_clipmask(src: Act, x: Optional[torch.Tensor], noise: Optional[torch.Tensor], clip: bool) -> torch.Tensor:
    dst = torch.empty((src.B, src.C, src.H, src.W), dtype=torch.float32, device=src.device)
    _chk("octave_nhwc_to_nchw_clipmask", lib.octave_nhwc_to_nchw_clipmask(_ref(src), _p(x), _p(noise), int(clip), dst.data_ptr(), stream_ptr()))
    return dst


def copy_window(src: Act, dst: Act, accumulate: bool = False) -> None:
    _chk("octave_copy_window", lib.octave_copy_window(_ref(src), _ref(dst), int(accumulate), stream_ptr()))


def space_to_depth(src: Act, H: int, W: int, want_chan_sum: bool = False):
    """-> dst, or (dst, per-channel fp64 sum of src) with want_chan_sum"""
    dst = Act.empty(src.B, H, W, 4 * src.C, src.dtype, src.device)
    G = src.C // 8
    fused = want_chan_sum and G <= 256 and 256 % G == 0
    cs = zeros_f64(src.C, src.device) if fused else None
    _chk("octave_space_to_depth", lib.octave_space_to_depth(_ref(src), _ref(dst), _p(cs), stream_ptr()))
    if want_chan_sum:
        return dst, (cs if fused else chan_stats(src)[:src.C])
    return dst


# --- convolutions ------------------------------------------------------------------------------------------
class ConvSpec:
    """Geometry + parameters of one Conv2d / ConvTranspose2d(k2,s2) and the cache of its bf16 operand packs."""

    def __init__(self, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter], cin: int, cout: int, k: int,
                 stride: int = 1, pad: int = 0, groups: int = 1, transposed: bool = False):
        self.weight, self.bias = weight, bias
        self.cin, self.cout, self.k, self.stride, self.pad, self.groups = cin, cout, k, stride, pad, groups
        self.transposed = transposed
        self._packs = {}
        # merge tiny groups into block-diagonal dense groups until a group fills a UMMA K/N chunk
        m = 1
        if not transposed:
            cg, og = cin // groups, cout // groups
            while (cg * m) % 32 or (og * m) % 32:
                m *= 2
                if m > groups or groups % m:
                    m = 0
                    break
        self.dense_groups = groups // m if m else 0

    def tc_ok(self, dtype: torch.dtype) -> bool:
        if dtype != torch.bfloat16:
            return False
        if self.transposed:
            return self.cin % 32 == 0 and self.cout % 32 == 0
        return self.dense_groups > 0 and self.stride == 1 and self.k in (1, 3) and self.pad == self.k // 2

    def _pack_key(self, mode: int):
        w = self.weight
        return (mode, w._version, w.data_ptr())

    def _pack_numel(self, mode: int) -> int:
        taps = self.k * self.k
        dg = max(self.dense_groups, 1)
        if mode == _lib_pack.FWD:
            return taps * self.cout * (self.cin // dg)
        if mode == _lib_pack.DGRAD:
            return taps * self.cin * (self.cout // dg)
        return 4 * self.cout * self.cin

    def pack_modes(self):
        """operand packs the tensor-core kernels of this conv consume (forward, data gradient)"""
        return (_lib_pack.CONVT_FWD, _lib_pack.CONVT_DGRAD) if self.transposed else (_lib_pack.FWD, _lib_pack.DGRAD)

    def pack(self, mode: int) -> torch.Tensor:
        key = self._pack_key(mode)
        hit = self._packs.get(mode)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = self.weight
        out = hit[1] if hit is not None and hit[1].device == w.device else \
            torch.empty(self._pack_numel(mode), dtype=torch.bfloat16, device=w.device)
        wd = w.detach()
        if wd.dtype != torch.float32 or not wd.is_contiguous():
            wd = wd.float().contiguous()
        _chk("octave_pack_weight", lib.octave_pack_weight(wd.data_ptr(), mode, self.cout, self.cin, self.groups,
                                                          max(self.dense_groups, 1), self.k, out.data_ptr(), stream_ptr()))
        self._packs[mode] = (key, out)
        return out


class MultiPacker:
    """Re-packs every stale bf16 operand of a list of ConvSpecs in ONE kernel launch (the per-step re-pack after the
    optimiser update).  The device job table is rebuilt only when a weight or pack buffer moves."""

    def __init__(self, specs):
        self.specs = [s for s in specs if s.weight.dtype == torch.float32]
        self._sig = None
        self._table = None
        self._n = 0
        self._blocks = 0

    def run(self) -> None:
        specs = self.specs
        if not specs:
            return
        stale = False
        for s in specs:
            for m in s.pack_modes():
                hit = s._packs.get(m)
                if hit is None or hit[0] != s._pack_key(m):
                    stale = True
                    break
            if stale:
                break
        if not stale:
            return
        dev = specs[0].weight.device
        jobs, sig, start = [], [], 0
        for s in specs:
            w = s.weight
            if not w.is_contiguous():
                raise RuntimeError("octave_b200: conv weights must be contiguous")
            for m in s.pack_modes():
                hit = s._packs.get(m)
                out = hit[1] if hit is not None and hit[1].device == dev else \
                    torch.empty(s._pack_numel(m), dtype=torch.bfloat16, device=dev)
                s._packs[m] = (s._pack_key(m), out)
                dg = max(s.dense_groups, 1)
                nb = lib.octave_pack_job_blocks(m, s.cout, s.cin, dg, s.k)
                jobs.append((w.data_ptr(), out.data_ptr(), m, s.cout, s.cin, s.groups, dg, s.k, start))
                sig.append((w.data_ptr(), out.data_ptr()))
                start += nb
        if sig != self._sig:
            arr = (PackJob * len(jobs))(*[PackJob(*j) for j in jobs])
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            self._table = host.to(dev)
            self._sig, self._n, self._blocks = sig, len(jobs), start
        _chk("octave_pack_weight_multi", lib.octave_pack_weight_multi(self._table.data_ptr(), self._n, self._blocks, stream_ptr()))


class _lib_pack:
    FWD, DGRAD, CONVT_FWD, CONVT_DGRAD = 0, 1, 2, 3


def _conv_desc(B, H, W, cin, cout, groups, k, stride, pad, x: Act, y: Act, Hout, Wout, mode=0, act=0, accumulate=False,
               real_groups=0) -> _lib.ConvDesc:
    d = _lib.ConvDesc()
    d.B, d.H, d.W, d.cin, d.cout, d.groups, d.ksize, d.stride, d.pad = B, H, W, cin, cout, groups, k, stride, pad
    d.x_ld, d.x_coff, d.y_ld, d.y_coff = x.ld, x.coff, y.ld, y.coff
    d.Hout, d.Wout, d.mode, d.relu = Hout, Wout, mode, act
    d.in_dtype, d.out_dtype = _dt(x.dtype), _dt(y.dtype)
    d.accumulate, d.real_groups = int(accumulate), real_groups
    return d


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def conv_out_hw(spec: ConvSpec, H: int, W: int) -> Tuple[int, int]:
    if spec.transposed:
        return 2 * H, 2 * W
    return (H + 2 * spec.pad - spec.k) // spec.stride + 1, (W + 2 * spec.pad - spec.k) // spec.stride + 1


def conv_fwd(x: Act, spec: ConvSpec, out: Optional[Act] = None, act: int = 0, want_stats: Optional[bool] = None):
    """Conv2d / ConvTranspose2d forward (+bias, + fused activation).  `out` may be a channel slice of a concat
    buffer and, for the transposed conv, may be spatially cropped.  want_stats (True/False, not None):
    return (out, stats) with stats = fp64 [2*cout] per-channel sum / sum of squares of the output when True (fused into
    the tensor-core epilogue when possible) and None when False."""
    Ho, Wo = conv_out_hw(spec, x.H, x.W)
    if out is None:
        out = Act.empty(x.B, Ho, Wo, spec.cout, x.dtype, x.device)
    bias = _f32(spec.bias)
    stats = None
    if spec.tc_ok(x.dtype) and act in (0, 1):
        if want_stats and not spec.transposed:
            stats = zeros_f64(2 * spec.cout, x.device)
        if spec.transposed:
            d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, 1, 1, 1, 0, x, out, out.H, out.W, mode=_lib.CONV_MODE_CONVT, act=act)
            wp = spec.pack(_lib_pack.CONVT_FWD)
        else:
            d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, spec.dense_groups, spec.k, 1, spec.k // 2, x, out, out.H, out.W, act=act)
            wp = spec.pack(_lib_pack.FWD)
        _chk("octave_conv_tc_fwd", lib.octave_conv_tc_fwd(C.byref(d), x.buf.data_ptr(), wp.data_ptr(), _p(bias), out.buf.data_ptr(), _p(stats), stream_ptr()))
        return out if want_stats is None else (out, stats)
    if spec.transposed:
        # CUDA-core path: 1x1 conv to [B,H,W,4*Cout] (+bias), then 2x2 pixel shuffle into the (cropped) output view
        w4 = _f32(spec.weight).permute(2, 3, 1, 0).reshape(4 * spec.cout, spec.cin, 1, 1).contiguous()
        b4 = bias.repeat(4) if bias is not None else None
        tmp = Act.empty(x.B, x.H, x.W, 4 * spec.cout, x.dtype, x.device)
        d = _conv_desc(x.B, x.H, x.W, spec.cin, 4 * spec.cout, 1, 1, 1, 0, x, tmp, x.H, x.W, act=act)
        _chk("octave_conv_direct_fwd", lib.octave_conv_direct_fwd(C.byref(d), x.buf.data_ptr(), w4.data_ptr(), _p(b4), tmp.buf.data_ptr(), stream_ptr()))
        _chk("octave_depth_to_space", lib.octave_depth_to_space(_ref(tmp), _ref(out), stream_ptr()))
        return out if want_stats is None else (out, chan_stats(out) if want_stats else None)
    d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, spec.groups, spec.k, spec.stride, spec.pad, x, out, out.H, out.W, act=act)
    _chk("octave_conv_direct_fwd", lib.octave_conv_direct_fwd(C.byref(d), x.buf.data_ptr(), _f32(spec.weight).data_ptr(), _p(bias),
                                                              out.buf.data_ptr(), stream_ptr()))
    return out if want_stats is None else (out, chan_stats(out) if want_stats else None)


def conv_dgrad(dy: Act, spec: ConvSpec, H: int, W: int, out: Optional[Act] = None, accumulate: bool = False) -> Act:
    """Gradient w.r.t. the conv input ([B,H,W,cin]).  dy is the gradient of the pre-activation output."""
    if out is None:
        out = Act.empty(dy.B, H, W, spec.cin, dy.dtype, dy.device)
    if spec.tc_ok(dy.dtype):
        # same implicit GEMM with cin/cout swapped and the flipped/transposed pack
        d = _conv_desc(dy.B, H, W, spec.cout, spec.cin, spec.dense_groups, spec.k, 1, spec.k // 2, dy, out, H, W, accumulate=accumulate)
        wp = spec.pack(_lib_pack.DGRAD)
        _chk("octave_conv_tc_fwd(dgrad)", lib.octave_conv_tc_fwd(C.byref(d), dy.buf.data_ptr(), wp.data_ptr(), None, out.buf.data_ptr(), None, stream_ptr()))
        return out
    d = _conv_desc(dy.B, H, W, spec.cin, spec.cout, spec.groups, spec.k, spec.stride, spec.pad, out, dy, dy.H, dy.W, accumulate=accumulate)
    _chk("octave_conv_direct_dgrad", lib.octave_conv_direct_dgrad(C.byref(d), dy.buf.data_ptr(), _f32(spec.weight).data_ptr(),
                                                                  out.buf.data_ptr(), stream_ptr()))
    return out


def conv_wgrad(x: Act, dy: Act, spec: ConvSpec, zero_bias_grad: bool = False):
    """-> (dW in the torch parameter layout fp32, dbias or None)"""
    dw = torch.empty(spec.weight.shape, dtype=torch.float32, device=x.device)
    db = None
    if zero_bias_grad and spec.bias is not None and spec.tc_ok(x.dtype):
        d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, spec.dense_groups, spec.k, 1, spec.k // 2, x, dy, dy.H, dy.W,
                       real_groups=spec.groups)
        _chk("octave_conv_tc_wgrad", lib.octave_conv_tc_wgrad(C.byref(d), x.buf.data_ptr(), dy.buf.data_ptr(), dw.data_ptr(), stream_ptr()))
        return dw, torch.zeros(spec.cout, dtype=torch.float32, device=x.device)
    if spec.tc_ok(x.dtype):
        d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, spec.dense_groups, spec.k, 1, spec.k // 2, x, dy, dy.H, dy.W,
                       real_groups=spec.groups)
        _chk("octave_conv_tc_wgrad", lib.octave_conv_tc_wgrad(C.byref(d), x.buf.data_ptr(), dy.buf.data_ptr(), dw.data_ptr(), stream_ptr()))
        if spec.bias is not None:
            db = chan_stats(dy)[:spec.cout].float()
        return dw, db
    if spec.bias is not None:
        db = torch.empty(spec.cout, dtype=torch.float32, device=x.device)
    d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, spec.groups, spec.k, spec.stride, spec.pad, x, dy, dy.H, dy.W)
    _chk("octave_conv_direct_wgrad", lib.octave_conv_direct_wgrad(C.byref(d), x.buf.data_ptr(), dy.buf.data_ptr(), dw.data_ptr(), _p(db), stream_ptr()))
    return dw, db


def convt_bwd(x: Act, dy: Act, spec: ConvSpec, need_dx: bool = True, side=None):
    """Backward of ConvTranspose2d(k=2,s=2): dy (possibly cropped view [B,Ho,Wo,cout]) -> (dx, dW [Cin,Cout,2,2], dbias).
    side: optional `run(fn, *acts)` that launches the weight gradient on another stream (network.on_side_stream)."""
    # [B,H,W,4*cout], zeros where dy was cropped; the same pass yields the bias gradient sum_pixels dy
    dys, dy_sum = space_to_depth(dy, x.H, x.W, want_chan_sum=spec.bias is not None) if spec.bias is not None else (space_to_depth(dy, x.H, x.W), None)
    if not spec.tc_ok(x.dtype):
        w4 = _f32(spec.weight).permute(2, 3, 1, 0).reshape(4 * spec.cout, spec.cin, 1, 1).contiguous()
        dw4 = torch.empty_like(w4)
        db4 = torch.empty(4 * spec.cout, dtype=torch.float32, device=x.device) if spec.bias is not None else None
        d = _conv_desc(x.B, x.H, x.W, spec.cin, 4 * spec.cout, 1, 1, 1, 0, x, dys, x.H, x.W)
        _chk("octave_conv_direct_wgrad", lib.octave_conv_direct_wgrad(C.byref(d), x.buf.data_ptr(), dys.buf.data_ptr(), dw4.data_ptr(), _p(db4), stream_ptr()))
        dw = dw4.reshape(2, 2, spec.cout, spec.cin).permute(3, 2, 0, 1).contiguous()
        db = db4.reshape(4, spec.cout).sum(0) if db4 is not None else None
        dx = None
        if need_dx:
            dx = Act.empty(x.B, x.H, x.W, spec.cin, x.dtype, x.device)
            _chk("octave_conv_direct_dgrad", lib.octave_conv_direct_dgrad(C.byref(_conv_desc(x.B, x.H, x.W, spec.cin, 4 * spec.cout, 1, 1, 1, 0, dx, dys, x.H, x.W)),
                                                                          dys.buf.data_ptr(), w4.data_ptr(), dx.buf.data_ptr(), stream_ptr()))
        return dx, dw, db
    def wgrad():
        dw_ = torch.empty(spec.weight.shape, dtype=torch.float32, device=x.device)
        d = _conv_desc(x.B, x.H, x.W, spec.cin, spec.cout, 1, 1, 1, 0, x, dys, x.H, x.W, mode=_lib.CONV_MODE_CONVT)
        _chk("octave_conv_tc_wgrad(convT)", lib.octave_conv_tc_wgrad(C.byref(d), x.buf.data_ptr(), dys.buf.data_ptr(), dw_.data_ptr(), stream_ptr()))
        return dw_

    dw = side(wgrad, x, dys) if side is not None else wgrad()
    db = dy_sum.float() if spec.bias is not None else None
    dx = None
    if need_dx:
        dx = Act.empty(x.B, x.H, x.W, spec.cin, x.dtype, x.device)
        dd = _conv_desc(x.B, x.H, x.W, 4 * spec.cout, spec.cin, 1, 1, 1, 0, dys, dx, x.H, x.W)
        wp = spec.pack(_lib_pack.CONVT_DGRAD)
        _chk("octave_conv_tc_fwd(convT dgrad)", lib.octave_conv_tc_fwd(C.byref(dd), dys.buf.data_ptr(), wp.data_ptr(), None, dx.buf.data_ptr(), None, stream_ptr()))
    return dx, dw, db


# --- discriminator: space-to-depth formulation on the tensor cores ------------------------------------------
def nchw_to_s2d(src: torch.Tensor, dst: Act, qs: int, coff: int, noise: Optional[torch.Tensor] = None, clip: bool = False) -> None:
    s = src.contiguous().float()
    B, C_, H, W = s.shape
    _chk("octave_nchw_to_s2d", lib.octave_nchw_to_s2d(s.data_ptr(), B, C_, H, W, _p(noise), int(clip), _ref(dst), qs, coff, stream_ptr()))


def s2d_to_nchw(src: Act, qs: int, coff: int, C_: int, H: int, W: int, x: Optional[torch.Tensor] = None,
                noise: Optional[torch.Tensor] = None, clip: bool = False) -> torch.Tensor:
    dst = torch.empty((src.B, C_, H, W), dtype=torch.float32, device=src.device)
    _chk("octave_s2d_to_nchw", lib.octave_s2d_to_nchw(_ref(src), qs, coff, C_, H, W, _p(x), _p(noise), int(clip), dst.data_ptr(), stream_ptr()))
    return dst


def pack_weight_s2d(w: torch.Tensor, scale: Optional[torch.Tensor], mode: int, qs: int) -> torch.Tensor:
    cout, cin, ksize = w.shape[0], w.shape[1], w.shape[2]
    out = torch.empty(9 * cout * 4 * qs, dtype=torch.bfloat16, device=w.device)
    _chk("octave_pack_weight_s2d", lib.octave_pack_weight_s2d(_f32(w).data_ptr(), _p(scale), mode, cout, cin, qs, ksize, out.data_ptr(), stream_ptr()))
    return out


def conv4x4s2_tc_fwd(xs: Act, wpack: torch.Tensor, bias: Optional[torch.Tensor], cout: int, Ho: int, Wo: int, act: int,
                     stats: Optional[torch.Tensor] = None) -> Act:
    """4x4 stride-2 pad-1 conv as a 3x3 conv over the space-to-depth input `xs` ([B,hs,ws,4*qs]); output [B,Ho,Wo,cout]."""
    y = Act.empty(xs.B, Ho, Wo, cout, xs.dtype, xs.device)
    d = _conv_desc(xs.B, xs.H, xs.W, xs.C, cout, 1, 3, 1, 1, xs, y, Ho, Wo, act=act)
    _chk("octave_conv_tc_fwd(s2d)", lib.octave_conv_tc_fwd(C.byref(d), xs.buf.data_ptr(), wpack.data_ptr(), _p(bias), y.buf.data_ptr(), _p(stats), stream_ptr()))
    return y


def conv4x4s2_tc_dgrad(dz: Act, wpack_d: torch.Tensor, hs: int, ws: int, K: int) -> Act:
    """Gradient w.r.t. the space-to-depth input: 3x3 conv over dz ([B,Ho,Wo,cout]) with the flipped pack -> [B,hs,ws,K]."""
    dx = Act.empty(dz.B, hs, ws, K, dz.dtype, dz.device)
    d = _conv_desc(dz.B, dz.H, dz.W, dz.C, K, 1, 3, 1, 1, dz, dx, hs, ws)
    _chk("octave_conv_tc_fwd(s2d dgrad)", lib.octave_conv_tc_fwd(C.byref(d), dz.buf.data_ptr(), wpack_d.data_ptr(), None, dx.buf.data_ptr(), None, stream_ptr()))
    return dx


def conv4x4s2_tc_wgrad(xs: Act, dz: Act, cin: int, qs: int, ksize: int = 4) -> torch.Tensor:
    """-> dW fp32 [cout][cin][k][k]"""
    cout = dz.C
    dw3 = torch.empty((cout, xs.C, 3, 3), dtype=torch.float32, device=xs.device)
    d = _conv_desc(xs.B, xs.H, xs.W, xs.C, cout, 1, 3, 1, 1, xs, dz, dz.H, dz.W)
    _chk("octave_conv_tc_wgrad(s2d)", lib.octave_conv_tc_wgrad(C.byref(d), xs.buf.data_ptr(), dz.buf.data_ptr(), dw3.data_ptr(), stream_ptr()))
    dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=xs.device)
    _chk("octave_unpack_wgrad_s2d", lib.octave_unpack_wgrad_s2d(dw3.data_ptr(), cout, cin, qs, ksize, dw.data_ptr(), stream_ptr()))
    return dw


def conv1x1_tc_s2d_store(x: Act, wpack: torch.Tensor, bias: torch.Tensor, cout: int, dst: Act, qs: int, act: int) -> None:
    """1x1 conv (+bias, +act) whose output pixel (h,w) is stored space-to-depth into `dst` ([B,ceil(h/2),ceil(w/2),4*qs])."""
    d = _conv_desc(x.B, x.H, x.W, x.C, cout, 1, 1, 1, 0, x, dst, x.H, x.W, act=act)
    d.out_s2d_qs = qs
    _chk("octave_conv_tc_fwd(s2d store)", lib.octave_conv_tc_fwd(C.byref(d), x.buf.data_ptr(), wpack.data_ptr(), bias.data_ptr(), dst.buf.data_ptr(), None, stream_ptr()))


def rowdot_fwd(x: Act, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty((x.B, 1), dtype=torch.float32, device=x.device)
    _chk("octave_rowdot_fwd", lib.octave_rowdot_fwd(_ref(x), w.data_ptr(), _p(bias), out.data_ptr(), stream_ptr()))
    return out


def rowdot_bwd(x: Act, w: torch.Tensor, g: torch.Tensor, need_dw: bool):
    dx = x.like()
    dw = torch.empty_like(w) if need_dw else None
    db = torch.empty(1, dtype=torch.float32, device=x.device) if need_dw else None
    _chk("octave_rowdot_bwd", lib.octave_rowdot_bwd(_ref(x), w.data_ptr(), g.data_ptr(), _ref(dx), _p(dw), _p(db), stream_ptr()))
    return dx, dw, db


def spectral_sigma(w2d: torch.Tensor, u: torch.Tensor, v: torch.Tensor, training: bool, eps: float = 1e-12) -> torch.Tensor:
    """One power iteration (training: u, v updated in place) and sigma = u . (W v) -> float32[2] = (sigma, 1/sigma)."""
    out = torch.empty(2, dtype=torch.float32, device=w2d.device)
    _chk("octave_spectral_sigma", lib.octave_spectral_sigma(w2d.data_ptr(), w2d.shape[0], w2d.shape[1], u.data_ptr(), v.data_ptr(),
                                                            int(training), eps, out.data_ptr(), stream_ptr()))
    return out


def spectral_sigma_multi(ws, us, vs, training: bool, eps: float = 1e-12):
    """The power iteration + sigma of every spectral-norm layer of one critic call in ONE launch.  ws: [rows, cols] fp32
    matrices; us / vs: the layers' weight_u / weight_v (updated in place when training).  -> per layer a float32 tensor
    [2 + rows + cols] = sigma, 1/sigma, snapshot of u, snapshot of v (what the backward pass of THIS call needs)."""
    n = len(ws)
    assert 0 < n <= SN_MAX_JOBS
    sizes = [2 + w.shape[0] + w.shape[1] for w in ws]
    offs = [0]
    for sz in sizes:
        offs.append(offs[-1] + ((sz + 3) & ~3))
    buf = torch.empty(offs[-1], dtype=torch.float32, device=ws[0].device)
    outs = [buf[offs[i]:offs[i] + sizes[i]] for i in range(n)]
    jobs = (OctaveSnJob * n)()
    for i in range(n):
        jobs[i] = OctaveSnJob(ws[i].data_ptr(), us[i].data_ptr(), vs[i].data_ptr(), outs[i].data_ptr(), ws[i].shape[0], ws[i].shape[1])
    _chk("octave_spectral_sigma_multi", lib.octave_spectral_sigma_multi(jobs, n, int(training), eps, stream_ptr()))
    return outs


def spectral_wgrad(dw: torch.Tensor, w_orig: torch.Tensor, u: torch.Tensor, v: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """dW_orig = dW / sigma - <dW, W_orig> / sigma^2 * u v^T (spectral norm backward with u, v constant); fp32, shape of W_orig."""
    rows = w_orig.shape[0]
    cols = w_orig.numel() // rows
    out = torch.empty_like(w_orig)
    parts = torch.empty(64, dtype=torch.float32, device=dw.device)
    _chk("octave_spectral_wgrad", lib.octave_spectral_wgrad(dw.data_ptr(), w_orig.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
                                                            rows, cols, parts.data_ptr(), out.data_ptr(), 0, stream_ptr()))
    return out
