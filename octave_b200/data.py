"""On-GPU input pipeline: synthetic OCTA-like batches, unpaired "real" mask pyramids and augmentation are
produced by CUDA kernels straight into device buffers (no host generation, no H2D), so a multi-GPU run is not input-bound.
The reference ships no data loader (README.md:39-47); the synthetic model is that of SURVEY.md §8d / octave_b200/synth.py."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import lib

AUG_FLIP_H, AUG_FLIP_V, AUG_ROT90, AUG_PHOTO = 1, 2, 4, 8
_vp = C.c_void_p
lib.octave_synth_octa.restype = C.c_int
lib.octave_synth_octa.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]
lib.octave_synth_mask_pyramid.restype = C.c_int
lib.octave_synth_mask_pyramid.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp), _vp]
lib.octave_augment.restype = C.c_int
lib.octave_augment.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp]


def _stream() -> int:
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class OnDeviceOcta:
    """Seeded stream of synthetic training batches on the GPU.

    next() -> (x [B,3,H,W], ys [B,2,H,W], real pyramid [5 x [B,2,H>>k,W>>k]]), fp32 device tensors that are overwritten by the
    following call (static buffers: they can be the inputs of a captured CUDA graph).  Batch i is a pure function of
    (seed, i, rank): different ranks draw different data (weak scaling), a re-run draws the same."""

    def __init__(self, B: int, H: int, W: int, device, seed: int = 0, rank: int = 0, n_ridges: int = 12, levels: int = 5,
                 augment: int = AUG_FLIP_H | AUG_FLIP_V | AUG_ROT90 | AUG_PHOTO):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("octave_b200.data: the input pipeline runs on the GPU only (no CPU fallback)")
        self.B, self.H, self.W, self.levels, self.n_ridges = B, H, W, levels, n_ridges
        self.seed, self.rank, self.augment, self.step = seed, rank, augment, 0
        self._raw_x = torch.empty(B, 3, H, W, device=dev)
        self._raw_ys = torch.empty(B, 2, H, W, device=dev)
        self.x = torch.empty_like(self._raw_x)
        self.ys = torch.empty_like(self._raw_ys)
        self.vessel = torch.empty(B, H, W, dtype=torch.uint8, device=dev)
        self.real = [torch.empty(B, 2, len(range(0, H, 2 ** k)), len(range(0, W, 2 ** k)), device=dev) for k in range(levels)]
        self._real_ptrs = (_vp * levels)(*[r.data_ptr() for r in self.real])

    def _seed(self, salt: int) -> int:
        return (self.seed * 0x9E3779B97F4A7C15 + self.rank * 0xD1B54A32D192ED03 + self.step * 0x2545F4914F6CDD1D + salt) & 0xFFFFFFFFFFFFFFFF

    def next(self) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor]]:
        s = _stream()
        _lib.check("octave_synth_octa", lib.octave_synth_octa(self._seed(1), self.B, self.H, self.W, self.n_ridges, self._raw_x.data_ptr(),
                                                              self._raw_ys.data_ptr(), self.vessel.data_ptr(), s))
        if self.augment:
            _lib.check("octave_augment", lib.octave_augment(self._seed(2), self.B, 3, 2, self.H, self.W, self.augment, self._raw_x.data_ptr(),
                                                            self._raw_ys.data_ptr(), self.x.data_ptr(), self.ys.data_ptr(), s))
        else:
            self.x.copy_(self._raw_x); self.ys.copy_(self._raw_ys)
        _lib.check("octave_synth_mask_pyramid", lib.octave_synth_mask_pyramid(self._seed(3), self.B, self.H, self.W, self.n_ridges, self.levels,
                                                                              self._real_ptrs, s))
        self.step += 1
        return self.x, self.ys, self.real


def augment(x: torch.Tensor, ys: torch.Tensor, seed: int, flags: int = AUG_FLIP_H | AUG_FLIP_V | AUG_ROT90 | AUG_PHOTO):
    """Augment any fp32 NCHW image batch `x` and its label maps `ys` with the same per-sample geometric transform."""
    if not (x.is_cuda and ys.is_cuda):
        raise RuntimeError("octave_b200.data.augment: inputs must be CUDA tensors (no CPU fallback)")
    x, ys = x.contiguous().float(), ys.contiguous().float()
    xo, yo = torch.empty_like(x), torch.empty_like(ys)
    B, Cx, H, W = x.shape
    _lib.check("octave_augment", lib.octave_augment(seed & 0xFFFFFFFFFFFFFFFF, B, Cx, ys.shape[1], H, W, flags, x.data_ptr(), ys.data_ptr(),
                                                    xo.data_ptr(), yo.data_ptr(), _stream()))
    return xo, yo
