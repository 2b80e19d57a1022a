"""Multi-tensor optimisers on the sm_100a kernel `octave_optim_multi` (include/octave_b200.h): the whole parameter list is
updated by ONE launch instead of torch's ~25 foreach kernels per step.  Update rules and state layout follow
torch.optim.SGD (momentum, weight_decay) and torch.optim.AdamW; `state_dict()` / `load_state_dict()` interchange with them.
The reference ships no optimiser (its training script is on an unmounted branch, README.md:39-47)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from . import _lib
from ._lib import lib


class OptJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("block_start", C.c_int64)]


class OptHyper(C.Structure):
    _fields_ = [("algo", C.c_int32), ("first_step", C.c_int32), ("lr", C.c_float), ("momentum", C.c_float), ("weight_decay", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("bias_correction1", C.c_float), ("bias_correction2", C.c_float)]


lib.octave_optim_job_blocks.restype = C.c_int64
lib.octave_optim_job_blocks.argtypes = [C.c_int64]
lib.octave_optim_multi.restype = C.c_int
lib.octave_optim_multi.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(OptHyper), C.c_void_p]


def _stream_ptr() -> int:
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _MultiTensorOptimizer:
    ALGO = 0

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("octave_b200.optim: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
        self.lr = lr
        self.steps = 0
        self.m = [None] * len(self.params)
        self.v = [None] * len(self.params)
        self._sig = None
        self._table = None
        self._blocks = 0
        self._n = 0
        self._ring, self._pin, self._dev, self._ev, self._captured = 0, [None, None], [None, None], [None, None], []

    # torch.optim interface ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _hyper(self, first: bool) -> OptHyper:
        raise NotImplementedError

    def _needs_v(self) -> bool:
        return False

    @torch.no_grad()
    def step(self) -> None:
        live = [(i, p) for i, p in enumerate(self.params) if p.grad is not None]
        if not live:
            return
        dev = live[0][1].device
        jobs, sig, start = [], [], 0
        for i, p in live:
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
                p.grad = g
            if self.m[i] is None:
                self.m[i] = torch.zeros_like(p)
            if self._needs_v() and self.v[i] is None:
                self.v[i] = torch.zeros_like(p)
            v = self.v[i].data_ptr() if self.v[i] is not None else 0
            jobs.append((p.data_ptr(), g.data_ptr(), self.m[i].data_ptr(), v, p.numel(), start))
            sig.append((p.data_ptr(), g.data_ptr()))
            start += lib.octave_optim_job_blocks(p.numel())
        if sig != self._sig:
            # (eager mode: gradients are fresh tensors every step, so the table is re-uploaded every step)
            arr = (OptJob * len(jobs))(*[OptJob(*j) for j in jobs])
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            if torch.cuda.is_current_stream_capturing():
                # the upload becomes a node of the graph and runs at every replay: its pinned source is never reused
                pin = raw.clone().pin_memory()
                self._captured.append(pin)
                self._table = torch.empty(raw.numel(), dtype=torch.uint8, device=dev)
                self._captured.append(self._table)
            else:
                r = self._ring = (self._ring + 1) % 2
                if self._pin[r] is None or self._pin[r].numel() < raw.numel():
                    self._pin[r] = torch.empty(max(raw.numel(), 1 << 16), dtype=torch.uint8).pin_memory()
                    self._dev[r] = torch.empty(self._pin[r].numel(), dtype=torch.uint8, device=dev)
                elif self._ev[r] is not None:
                    self._ev[r].synchronize()      # the upload issued two steps ago from this slot has run
                pin = self._pin[r][:raw.numel()]
                pin.copy_(raw)
                self._table = self._dev[r][:raw.numel()]
            self._table.copy_(pin, non_blocking=True)
            if not torch.cuda.is_current_stream_capturing():
                self._ev[self._ring] = self._ev[self._ring] or torch.cuda.Event()
                self._ev[self._ring].record()
            self._sig, self._n, self._blocks = sig, len(jobs), start
        h = self._hyper(self.steps == 0)
        _lib.check("octave_optim_multi", lib.octave_optim_multi(self._table.data_ptr(), self._n, self._blocks, C.byref(h), _stream_ptr()))
        # the kernel wrote the parameters behind torch's back: bump their version counters, which is what invalidates the
        # cached bf16 operand packs (ops.ConvSpec) and folded inference weights
        torch.autograd.graph.increment_version([p for _, p in live])
        self.steps += 1

    # state interchange with torch.optim ---------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"steps": self.steps, "lr": self.lr, "m": self.m, "v": self.v}

    def load_state_dict(self, sd: dict) -> None:
        self.steps, self.lr = sd["steps"], sd["lr"]
        self.m, self.v = list(sd["m"]), list(sd["v"])
        self._sig = None


class FusedSGD(_MultiTensorOptimizer):
    """torch.optim.SGD(params, lr, momentum, weight_decay) semantics (dampening 0, no Nesterov) in one launch per step."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0):
        super().__init__(params, lr)
        self.momentum, self.weight_decay = momentum, weight_decay

    def _hyper(self, first: bool) -> OptHyper:
        return OptHyper(0, int(first), self.lr, self.momentum, self.weight_decay, 0.0, 0.0, 0.0, 1.0, 1.0)

    @property
    def momentum_buffers(self):
        return self.m


class FusedAdamW(_MultiTensorOptimizer):
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay) semantics (no amsgrad) in one launch per step.  One step
    count serves every parameter's bias corrections (torch counts per parameter: the two differ only for a parameter that
    had no gradient in some step).  The bias corrections are host values: re-capture a CUDA graph that contains step()."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, lr)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay

    def _needs_v(self) -> bool:
        return True

    def _hyper(self, first: bool) -> OptHyper:
        t = self.steps + 1
        return OptHyper(1, int(first), self.lr, 0.0, self.weight_decay, self.betas[0], self.betas[1], self.eps,
                        1.0 - self.betas[0] ** t, 1.0 - self.betas[1] ** t)
