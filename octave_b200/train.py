"""Scribble-supervised adversarial training step composed from the reference-named parts (the
reference ships no loop — README.md:35 — so the composition follows docs/figure-1 of the reference):

  G-step: att, agg, _ = seg(x);  L = WPCE(softmax(agg), scribble) + l_kl * KLD(att) + l_g * LSG(D(att));  backward; step
  D-step: L = LSD(D(real pyramid), D(att.detach()));  backward; step

Data parallelism (one process per GPU): gradients are averaged with bucketed NCCL all-reduces issued on a side stream
as soon as a bucket's gradients exist (the segmentor's explicit backward announces them stage by stage, decoder first),
so the exchange overlaps the rest of the backward pass.  BatchNorm statistics and loss normalisers stay per replica,
exactly like the reference modules under plain DDP (there is no SyncBN in the reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn, Tensor

from .losses import FusedSegmentorLoss, LSDiscriminatorialLoss
from .model import OctaScribbleNet


class GradAllReducer:
    """Bucketed gradient averaging over NCCL/gloo.  `reduce(params, grads)` may be called several times per backward
    with disjoint lists (in the order the gradients become ready): every ~bucket_bytes the gradients gathered so far are
    packed into ONE flat buffer (a single concat kernel) whose all-reduce starts at once on a side stream, overlapping the
    rest of the backward pass.  `finish()` waits for the buckets and re-points each `param.grad` at its slice of the
    averaged flat buffer — no copy back, no per-tensor kernels.

    grad_dtype="bf16" (NCCL on CUDA only): a bucket is gathered AND cast to bf16 by one kernel (octave_grad_pack_bf16), the
    all-reduce moves half the bytes, and one kernel writes the averaged values back into the fp32 gradient tensors
    (octave_grad_unpack_bf16).  The averaged gradient then carries one bf16 rounding of each rank's contribution plus NCCL's
    bf16 accumulation (relative error ~2^-8): the mode of the bf16 product path, whose gradients are bf16-accurate anyway."""

    def __init__(self, bucket_bytes: int = 25 << 20, process_group=None, grad_dtype: str = "fp32"):
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError(f"grad_dtype must be 'fp32' or 'bf16', got {grad_dtype!r}")
        self.grad_dtype = grad_dtype
        self.bucket_bytes = bucket_bytes
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._pending = []          # (flat, params, grads, work)
        self._cur_p: List[Optional[nn.Parameter]] = []
        self._cur_g: List[Tensor] = []
        self._cur_bytes = 0
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        backend = dist.get_backend(process_group) if dist.is_initialized() else ""
        self._avg = backend == "nccl"      # ncclAvg; gloo has no AVG: sum, then scale

    def _flush(self):
        if not self._cur_g:
            return
        params, grads = self._cur_p, self._cur_g
        self._cur_p, self._cur_g, self._cur_bytes = [], [], 0
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self._stream is not None and grads[0].is_cuda:
            self._stream.wait_stream(torch.cuda.current_stream())
            from . import network as _n
            side = _n._side_streams.get(grads[0].device)
            if side is not None:
                self._stream.wait_stream(side)       # weight gradients are produced on the side stream (config.overlap_wgrad)
            offs = None
            if (self.grad_dtype == "bf16" and self._avg
                    and all(g.dtype == torch.float32 and g.is_contiguous() and g.numel() < (1 << 31) for g in grads)):
                offs, off = [], 0
                for g in grads:
                    offs.append(off)
                    off += (g.numel() + 7) & ~7              # 16-byte slots; the padding is never read back
            with torch.cuda.stream(self._stream):
                if offs is not None:
                    flat = torch.empty(off, dtype=torch.bfloat16, device=grads[0].device)
                    self._bucket_kernel("octave_grad_pack_bf16", grads, offs, flat)
                else:
                    flat = torch.cat([g.reshape(-1) for g in grads])
                work = dist.all_reduce(flat, op=op, group=self.pg, async_op=True)
            for g in grads:
                g.record_stream(self._stream)
            self._pending.append((flat, params, grads, work, offs))
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=op, group=self.pg, async_op=True)
        self._pending.append((flat, params, grads, work, None))

    @staticmethod
    def _bucket_kernel(name: str, grads, offs, flat) -> None:
        """octave_grad_pack_bf16 / octave_grad_unpack_bf16 on the current stream, at most OCTAVE_GRAD_MAX_JOBS tensors per launch."""
        from . import _lib, ops
        fn = getattr(_lib.lib, name)
        for c0 in range(0, len(grads), _lib.GRAD_MAX_JOBS):
            chunk = list(zip(grads[c0:c0 + _lib.GRAD_MAX_JOBS], offs[c0:c0 + _lib.GRAD_MAX_JOBS]))
            arr = (_lib.OctaveGradJob * len(chunk))()
            blocks = 0
            for k, (g, o) in enumerate(chunk):
                arr[k] = _lib.OctaveGradJob(g.data_ptr(), o, g.numel(), blocks)
                blocks += _lib.lib.octave_optim_job_blocks(g.numel())
            _lib.check(name, fn(arr, len(chunk), flat.data_ptr(), ops.stream_ptr()))

    def reduce(self, params, grads: Optional[Sequence[Optional[Tensor]]] = None):
        """reduce(params, grads) — or reduce(grads): gradients without an owning parameter are averaged in place."""
        if self.world == 1:
            # nothing to exchange; the caller (the segmentor's autograd node with a hook installed) still expects the
            # reducer to own param.grad
            if grads is not None:
                for p, g in zip(params, grads):
                    if p is not None and g is not None:
                        p.grad = g.view_as(p)
            return
        if grads is None:
            params, grads = None, params
        if params is None:
            params = [None] * len(grads)
        for p, g in zip(params, grads):
            if g is None:
                continue
            self._cur_p.append(p)
            self._cur_g.append(g)
            self._cur_bytes += g.numel() * g.element_size()
            if self._cur_bytes >= self.bucket_bytes:
                self._flush()

    def finish(self):
        if self.world == 1:
            return
        self._flush()
        on_side = self._stream is not None and self._pending and self._pending[0][0].is_cuda
        for flat, params, grads, work, offs in self._pending:
            work.wait()
            ctxm = torch.cuda.stream(self._stream) if on_side else _null()
            with ctxm:
                if not self._avg:
                    flat.div_(self.world)
                if offs is not None:
                    # bf16 bucket: the averaged values go back into the fp32 gradient tensors (param.grad is unchanged)
                    self._bucket_kernel("octave_grad_unpack_bf16", grads, offs, flat)
            if offs is not None:
                for p, g in zip(params, grads):
                    if p is not None:
                        p.grad = g.view_as(p)
                continue
            off = 0
            for p, g in zip(params, grads):
                n = g.numel()
                view = flat[off:off + n].view(g.shape)
                if p is not None:
                    p.grad = view.view_as(p)          # the optimiser reads the averaged bucket slice directly
                else:
                    g.copy_(view)
                off += n
        if on_side:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._pending = []


class HostPrefetcher:
    """Input pipelining for the eager step: `put(x, ys, real)` (pinned host tensors) starts their upload on a copy stream
    while the current step computes; `get()` hands the device tensors to the compute stream."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self._batch = None

    def put(self, x: Tensor, ys: Tensor, real: Sequence[Tensor]) -> None:
        with torch.cuda.stream(self.stream):
            self._batch = (x.to(self.device, non_blocking=True), ys.to(self.device, non_blocking=True),
                           [r.to(self.device, non_blocking=True) for r in real])

    def get(self):
        if self._batch is None:
            return None
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream)
        x, ys, real = self._batch
        for t in (x, ys, *real):
            t.record_stream(cur)           # allocated on the copy stream, consumed on the compute stream
        self._batch = None
        return x, ys, real


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _momentum_buffers(opt) -> List[Tensor]:
    """momentum buffers that exist right now (torch.optim.SGD state or octave_b200.optim.FusedSGD.m)"""
    if hasattr(opt, "momentum_buffers"):
        return [b for b in opt.momentum_buffers if b is not None]
    return [st["momentum_buffer"] for st in opt.state.values() if st.get("momentum_buffer") is not None]


class TrainStep:
    """Owns the optimisers and runs one adversarial iteration.  `lambda_kl`, `lambda_g` and SGD(momentum) are this
    harness' choice (the mounted branch of the reference specifies none)."""

    def __init__(self, net: OctaScribbleNet, lr: float = 1e-3, momentum: float = 0.9, lambda_kl: float = 0.1,
                 lambda_g: float = 0.1, distributed: bool = False, bucket_bytes: int = 25 << 20, fused_optimizer: bool = True,
                 grad_dtype: Optional[str] = None):
        self.net = net
        self.lambda_kl, self.lambda_g = lambda_kl, lambda_g
        self.seg_params = [p for n, p in net.segmentor.named_parameters() if not n.startswith("linear_head_")]
        # one multi-tensor launch per optimiser step (octave_optim_multi) on the GPU; torch.optim.SGD otherwise (CPU tests)
        fused = fused_optimizer and all(p.is_cuda for p in self.seg_params)
        from .optim import FusedSGD
        make_opt = (lambda ps: FusedSGD(ps, lr=lr, momentum=momentum)) if fused else (lambda ps: torch.optim.SGD(ps, lr=lr, momentum=momentum))
        self.opt_g = make_opt(self.seg_params)
        self.has_d = hasattr(net, "discriminator")
        if self.has_d:
            self.dis_params = list(net.discriminator.parameters())
            self.opt_d = make_opt(self.dis_params)
        self.loss = FusedSegmentorLoss(weakly_supervise=isinstance(net.supervised_loss, nn.Module)
                                       and net.supervised_loss.__class__.__name__ == "WeightedPartialCE")
        self.lsd = LSDiscriminatorialLoss()
        if grad_dtype is None:
            # bf16 gradient buckets go with the bf16 product path (whose gradients are bf16-accurate anyway); fp32 mode keeps fp32
            from . import config
            grad_dtype = "bf16" if (config.compute_dtype == "bf16" and all(p.is_cuda for p in self.seg_params)) else "fp32"
        self.reducer = GradAllReducer(bucket_bytes, grad_dtype=grad_dtype) if distributed else None
        if self.reducer is not None:
            red = self.reducer
            net.segmentor._grad_ready_hook = lambda params, grads: red.reduce(params, grads)

    def _set_d_grad(self, flag: bool):
        if self.has_d:
            for p in self.dis_params:
                p.requires_grad_(flag)

    def g_step(self, x: Tensor, ys: Tensor, defer_join: bool = False) -> Dict[str, Tensor]:
        """defer_join (used by step()): the segmentor's optimiser step and the re-pack of its bf16 conv operands are left
        running on the side stream, beside the D-step that follows (which touches neither the segmentor's weights nor its
        gradients); step() joins before it returns."""
        net = self.net
        self.opt_g.zero_grad(set_to_none=True)
        self._set_d_grad(False)
        att, agg, _ = net.segmentor(x)
        y_fake = net.discriminator(att) if self.has_d else None
        # values and gradients of  sup + l_kl * KLD + l_g * LSG  from one sweep over the maps (K9 single pass)
        res = self.loss.total(agg, ys, att, y_fake, 1.0, self.lambda_kl, self.lambda_g)
        total = res['total']
        total.backward()
        if self.reducer is not None:
            self.reducer.finish()
        from . import config, network
        if config.overlap_wgrad and total.is_cuda and config.compute_dtype == "bf16":
            network.on_side_stream(self._update_segmentor)
            if not defer_join:
                network.join_side_stream()
        else:
            self._update_segmentor()
        self._set_d_grad(True)
        res['total'] = total.detach()
        res['attentions'] = att
        return res

    def _update_segmentor(self) -> None:
        self.opt_g.step()
        if hasattr(self.net.segmentor, "_repack"):
            self.net.segmentor._repack()         # bf16 operand packs of the updated conv weights (one launch)

    def d_step(self, real: Sequence[Tensor], fake: Sequence[Tensor]) -> Tensor:
        net = self.net
        self.opt_d.zero_grad(set_to_none=True)
        loss = self.lsd(net.discriminator(list(real)), net.discriminator([f.detach() for f in fake]))
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce(self.dis_params, [p.grad for p in self.dis_params])
            self.reducer.finish()
        self.opt_d.step()
        return loss.detach()

    def step(self, x: Tensor, ys: Tensor, real: Optional[Sequence[Tensor]] = None) -> Dict[str, Tensor]:
        res = self.g_step(x, ys, defer_join=True)
        if self.has_d and real is not None:
            res['discriminator'] = self.d_step(real, res['attentions'])
        from . import network
        network.join_side_stream()
        return res

    # ---- whole step as one CUDA graph ---------------------------------------------------------------------------
    def step_graphed(self, x: Tensor, ys: Tensor, real: Sequence[Tensor]) -> Dict[str, Tensor]:
        """Same arithmetic as `step`, replayed from one captured CUDA graph (forward, losses, backward, gradient
        all-reduce, optimiser steps, weight re-pack): ~1100 kernel launches cost one graph launch.  Inputs are copied into
        static device buffers (x / ys / real may live in pinned host memory); the returned tensors are static too and are
        overwritten by the next call.  The critic's per-call CPU randomness is drawn on the host before every replay in
        the eager order (discriminator.HostRandomFeed), so the CPU generator advances exactly as in eager mode.
        The learning rate, momentum and loss weights are baked into the captured graph (re-capture after changing them:
        `self._graph = None; self.graph_error = None`); batch shapes are static.
        Falls back to `step` (and says so in `self.graph_error`) when capture is not possible."""
        if getattr(self, "_graph", None) is None and getattr(self, "graph_error", None) is None:
            self._capture(x, ys, real)
        elif self._graph is not None and (tuple(x.shape) != tuple(self._sx.shape) or tuple(ys.shape) != tuple(self._sys.shape)
                                          or len(real) != len(self._sreal)):
            raise ValueError(f"octave_b200: step_graphed was captured for x {tuple(self._sx.shape)} / ys {tuple(self._sys.shape)}; "
                             f"got {tuple(x.shape)} / {tuple(ys.shape)} (a captured CUDA graph has static shapes)")
        if self._graph is None:
            dev = next(self.net.parameters()).device
            return self.step(x.to(dev, non_blocking=True), ys.to(dev, non_blocking=True), [r.to(dev, non_blocking=True) for r in real])
        pf = getattr(self, "_prefetched", None)
        if pf is not None and pf[0] is x and pf[1] is ys:
            # this batch was uploaded by prefetch() while the previous step computed: device-to-device hand-over
            torch.cuda.current_stream().wait_stream(self._copy_stream)
            src_x, src_ys, src_real = self._stage
            self._prefetched = None
        else:
            src_x, src_ys, src_real = x, ys, real
        self._sx.copy_(src_x, non_blocking=True)
        self._sys.copy_(src_ys, non_blocking=True)
        for d, r in zip(self._sreal, src_real):
            d.copy_(r, non_blocking=True)
        if getattr(self, "_copy_stream", None) is not None:
            self._handover.record()          # staging buffers are free again once these copies have run
        if self._feed is not None:
            # fresh CPU draws for this step: filled into the next pinned ring slot and uploaded on this stream BEFORE the
            # replay (outside the graph, so the pinned source rotates and the host never overwrites a pending upload)
            if not getattr(self, "_predrawn", False):
                self._feed.draw()
            self._predrawn = False
            self._feed.upload()
        self._graph.replay()
        return self._sout

    def predraw(self) -> None:
        """Make the NEXT step's CPU noise draws now (e.g. while the current replay is still running on the GPU and before the
        caller blocks on its result): the following step_graphed() then only uploads them.  Same generator, same order."""
        if getattr(self, "_graph", None) is not None and self._feed is not None and not getattr(self, "_predrawn", False):
            self._feed.draw()
            self._predrawn = True

    def prefetch(self, x: Tensor, ys: Tensor, real: Sequence[Tensor]) -> None:
        """Start the host-to-device upload of the NEXT batch (pinned host tensors) on a copy stream so that it overlaps
        the step that is computing now; the following `step_graphed(x, ys, real)` with the same tensors picks it up with a
        device-to-device copy.  No-op until the graph exists."""
        if getattr(self, "_graph", None) is None:
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
            self._handover = torch.cuda.Event()
            self._handover.record()
            self._stage = (torch.empty_like(self._sx), torch.empty_like(self._sys), [torch.empty_like(r) for r in self._sreal])
        cs = self._copy_stream
        cs.wait_event(self._handover)        # only the last hand-over reads the staging buffers — NOT the running replay
        with torch.cuda.stream(cs):
            self._stage[0].copy_(x, non_blocking=True)
            self._stage[1].copy_(ys, non_blocking=True)
            for d, r in zip(self._stage[2], real):
                d.copy_(r, non_blocking=True)
        self._prefetched = (x, ys)

    def _capture(self, x: Tensor, ys: Tensor, real: Sequence[Tensor]) -> None:
        from .discriminator import HostRandomFeed
        self._graph, self.graph_error, self._feed = None, None, None
        net = self.net
        dev = next(net.parameters()).device
        try:
            if not self.has_d or not net.discriminator._use_tc():
                raise RuntimeError("graph mode needs the bf16 tensor-core critic path")
            self._sx = x.to(dev).clone()
            self._sys = ys.to(dev).clone()
            self._sreal = [r.to(dev).clone() for r in real]
            self._feed = HostRandomFeed(net.discriminator, 3, dev)       # D(fake) in the G-step, D(real) + D(fake) in the D-step
            net.discriminator._rand_feed = self._feed
            # The warm-up steps (allocator, lazy state) must not train: parameters, BatchNorm / spectral-norm buffers and
            # the optimiser state are put back afterwards, so the first replay is the first optimiser step on this batch.
            snap = {k: v.detach().clone() for k, v in net.state_dict().items()}
            opts = [self.opt_g] + ([self.opt_d] if self.has_d else [])
            had_mom = [{id(b): b.detach().clone() for b in _momentum_buffers(o)} for o in opts]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                                        # warm-up on a side stream
                    self._feed.draw(); self._feed.upload()
                    self.step(self._sx, self._sys, self._sreal)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                cur = net.state_dict()
                for k, v in snap.items():
                    cur[k].copy_(v)
                for o, before in zip(opts, had_mom):
                    for buf in _momentum_buffers(o):
                        if id(buf) in before:
                            buf.copy_(before[id(buf)])
                        else:
                            buf.zero_()            # SGD: zeroed momentum + first gradient == a fresh first step
            del snap, had_mom
            from . import _lib
            g = torch.cuda.CUDAGraph()
            l0 = _lib.lib.octave_launch_count()             # (no host draw here: capture executes nothing)
            with torch.cuda.graph(g):
                self._feed.rewind()
                out = self.step(self._sx, self._sys, self._sreal)
            self.graph_kernel_nodes = int(_lib.lib.octave_launch_count() - l0)   # this library's kernels per replay
            out.pop('attentions', None)
            self._sout = out
            self._graph = g
        except Exception as e:  # capture is an optimisation: keep training eagerly, but say why
            self.graph_error = repr(e)[:500]
            self._graph = None
            if self.has_d and getattr(net.discriminator, "_rand_feed", None) is not None:
                net.discriminator._rand_feed = None
            self._feed = None
            torch.cuda.synchronize()
