"""Loss modules of the OCTAve training step, backed by the fused sm_100a loss kernel (K9).

Mirrors the reference interface (names, arguments, error behaviour):
  WeightedPartialCE, DiceLoss, InterlayerDivergence   /root/reference/architectures/segmentor/losses.py
  LSDiscriminatorialLoss, LSGeneratorLoss             /root/reference/architectures/discriminator/losses.py
plus `FusedSegmentorLoss`, which evaluates every term of a G-step in ONE statistics launch and ONE
gradient launch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn, Tensor

from . import _lib, config

_OUT_WPCE, _OUT_DICE, _OUT_KLD, _OUT_LSG, _OUT_LSD, _OUT_NAN = 0, 1, 2, 3, 4, 5


def _stream_ptr() -> int:
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"octave_b200: `{name}` is on {t.device}; the B200 kernels have no CPU fallback — move inputs to CUDA")


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ptr_array(ts: Sequence[Tensor]):
    arr = (C.c_void_p * 5)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


class _LossCfg:
    """Static (non-tensor) configuration of one fused-loss evaluation."""

    def __init__(self, flags: int, wpce_scale_mode: str = "mean", dice_eps: float = 1e-12,
                 att_weights: Sequence[float] = (), sum_weights: float = 1.0, jsd_eps: float = 1e-12):
        self.flags = flags
        self.wpce_scale_mode = wpce_scale_mode
        self.dice_eps = dice_eps
        self.att_weights = list(att_weights)
        self.sum_weights = sum_weights
        self.jsd_eps = jsd_eps


def _build_desc(cfg: _LossCfg, yhat: Optional[Tensor], att: Sequence[Tensor], d_real: Optional[Tensor],
                d_fake: Optional[Tensor]) -> _lib.LossDesc:
    d = _lib.LossDesc()
    ref = yhat if yhat is not None else (att[0] if att else None)
    if ref is not None:
        d.dtype = _lib.DTYPE_BF16 if ref.dtype == torch.bfloat16 else _lib.DTYPE_F32
        d.B, d.C, d.H, d.W = ref.shape
    else:
        d.dtype = _lib.DTYPE_F32
        d.B = d.C = d.H = d.W = 0
    d.flags = cfg.flags
    d.n_att = len(att)
    for k, a in enumerate(att):
        d.att_h[k], d.att_w[k] = a.shape[2], a.shape[3]
    for k, w in enumerate(cfg.att_weights[:4]):
        d.att_weight[k] = float(w)
    d.sum_weights = float(cfg.sum_weights)
    npix = max(d.B * d.H * d.W, 1)
    d.wpce_scale = 1.0 / npix if cfg.wpce_scale_mode == "mean" else 1.0
    d.dice_eps = float(cfg.dice_eps)
    d.jsd_eps = float(cfg.jsd_eps)
    d.n_real = 0 if d_real is None else d_real.numel()
    d.n_fake = 0 if d_fake is None else d_fake.numel()
    return d


class _FusedLossFn(torch.autograd.Function):
    """out[8] = (wpce, dice, kld, lsg, lsd, nanflag, 0, 0); one launch forward, one launch backward."""

    @staticmethod
    def forward(ctx, cfg: _LossCfg, yhat, ys, d_real, d_fake, *att):
        dev_ref = yhat if yhat is not None else (att[0] if att else d_fake)
        desc = _build_desc(cfg, yhat, att, d_real, d_fake)
        stats = torch.empty(_lib.lib.octave_loss_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev_ref.device)
        out = torch.empty(_lib.LOSS_OUT_SLOTS, dtype=torch.float32, device=dev_ref.device)
        att_arr = _ptr_array(att) if att else None
        rc = _lib.lib.octave_loss_fwd(C.byref(desc), _ptr(yhat), _ptr(ys), att_arr, _ptr(d_real), _ptr(d_fake),
                                      stats.data_ptr(), out.data_ptr(), _stream_ptr())
        _lib.check("octave_loss_fwd", rc)
        ctx.cfg, ctx.desc, ctx.n_att = cfg, desc, len(att)
        ctx.save_for_backward(stats, *(t for t in (yhat, ys, d_real, d_fake) if t is not None), *att)
        ctx.present = tuple(t is not None for t in (yhat, ys, d_real, d_fake))
        return out

    @staticmethod
    def backward(ctx, g_out):
        saved = list(ctx.saved_tensors)
        stats = saved.pop(0)
        vals = []
        for p in ctx.present:
            vals.append(saved.pop(0) if p else None)
        yhat, ys, d_real, d_fake = vals
        att = saved
        gscale = g_out.contiguous().float()
        flags = ctx.cfg.flags
        g_yhat = torch.empty_like(yhat) if (yhat is not None and flags & (_lib.LOSS_WPCE | _lib.LOSS_DICE)) else None
        g_att = [torch.empty_like(a) for a in att] if flags & _lib.LOSS_KLD else []
        g_real = torch.empty_like(d_real) if (d_real is not None and flags & _lib.LOSS_LSD) else None
        g_fake = torch.empty_like(d_fake) if (d_fake is not None and flags & (_lib.LOSS_LSG | _lib.LOSS_LSD)) else None
        rc = _lib.lib.octave_loss_bwd(C.byref(ctx.desc), _ptr(yhat), _ptr(ys), _ptr_array(att) if att else None,
                                      _ptr(d_real), _ptr(d_fake), stats.data_ptr(), gscale.data_ptr(), _ptr(g_yhat),
                                      _ptr_array(g_att) if g_att else None, _ptr(g_real), _ptr(g_fake), _stream_ptr())
        _lib.check("octave_loss_bwd", rc)
        return (None, g_yhat, None, g_real, g_fake, *(g_att if g_att else [None] * len(att)))


_fused_ws = {}


def _fused_workspace(dev, nbytes: int) -> Tensor:
    """Statistics workspace of the single-pass loss: zeroed once, left zero by every evaluation (the kernel cleans up after
    itself), one per (device, stream) because evaluations on one stream are serialised."""
    key = (dev, _stream_ptr())
    ws = _fused_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
        _fused_ws[key] = ws
    return ws


class _FusedTotalFn(torch.autograd.Function):
    """total = l_sup * WPCE + l_kl * KLD + l_g * LSG with values AND gradients from one sweep over the maps
    (octave_loss_fused: labels-only count pre-pass + one fused pass).  Returns (total, out[8]); `out` is not
    differentiable.  backward scales the stored gradients by the upstream gradient of `total` on the device
    (octave_loss_scale_grads, a no-op when it is 1)."""

    @staticmethod
    def forward(ctx, cfg: _LossCfg, lambdas, yhat, ys, d_fake, *att):
        desc = _build_desc(cfg, yhat, att, None, d_fake)
        dev = yhat.device
        stats = _fused_workspace(dev, _lib.lib.octave_loss_fused_stats_bytes(C.byref(desc)))
        out = torch.empty(_lib.LOSS_OUT_SLOTS, dtype=torch.float32, device=dev)
        g_yhat = torch.empty_like(yhat)
        g_att = [torch.empty_like(a) for a in att]
        g_fake = torch.empty_like(d_fake) if d_fake is not None else None
        lam = (C.c_float * 3)(*[float(v) for v in lambdas])
        rc = _lib.lib.octave_loss_fused(C.byref(desc), yhat.data_ptr(), ys.data_ptr(), _ptr_array(att), _ptr(d_fake), lam,
                                        stats.data_ptr(), out.data_ptr(), g_yhat.data_ptr(), _ptr_array(g_att), _ptr(g_fake),
                                        _stream_ptr())
        _lib.check("octave_loss_fused", rc)
        ctx.desc, ctx.n_att, ctx.has_fake = desc, len(att), d_fake is not None
        ctx.save_for_backward(g_yhat, *g_att, *([g_fake] if g_fake is not None else []))
        total = out[_lib.LOSS_OUT_TOTAL].clone()
        ctx.mark_non_differentiable(out)
        return total, out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        saved = list(ctx.saved_tensors)
        g_yhat, g_att = saved[0], saved[1:1 + ctx.n_att]
        g_fake = saved[1 + ctx.n_att] if ctx.has_fake else None
        g = g_total.reshape(1).float().contiguous()
        rc = _lib.lib.octave_loss_scale_grads(C.byref(ctx.desc), g.data_ptr(), g_yhat.data_ptr(), _ptr_array(g_att), _ptr(g_fake),
                                              _stream_ptr())
        _lib.check("octave_loss_scale_grads", rc)
        return (None, None, g_yhat, None, g_fake, *g_att)


def _prep_map(t: Tensor, name: str, dtype: Optional[torch.dtype] = None) -> Tensor:
    _require_cuda(t, name)
    if dtype is None:
        dtype = t.dtype if t.dtype in (torch.float32, torch.bfloat16) else torch.float32
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _prep_logits(t: Tensor, name: str) -> Tensor:
    _require_cuda(t, name)
    return t.contiguous().float()


def _common_dtype(*ts: Tensor) -> torch.dtype:
    return torch.bfloat16 if all(t.dtype == torch.bfloat16 for t in ts) else torch.float32


def fused_loss(cfg: _LossCfg, yhat=None, ys=None, att: Sequence[Tensor] = (), d_real=None, d_fake=None) -> Tensor:
    """Evaluate the enabled loss terms; returns the float32[8] output vector (see include/octave_b200.h)."""
    maps = [t for t in (yhat, *att) if t is not None]
    dtype = _common_dtype(*maps) if maps else torch.float32
    if yhat is not None:
        yhat = _prep_map(yhat, "y_hat", dtype)
        ys = _prep_map(ys, "ys", dtype)
    att = [_prep_map(a, f"attentions[{i}]", dtype) for i, a in enumerate(att)]
    if d_real is not None:
        d_real = _prep_logits(d_real, "y_real")
    if d_fake is not None:
        d_fake = _prep_logits(d_fake, "y_fake")
    return _FusedLossFn.apply(cfg, yhat, ys, d_real, d_fake, *att)


# -------------------------------------------------------------------------------------------------
# Reference-named modules
# -------------------------------------------------------------------------------------------------
class _WpceAltFn(torch.autograd.Function):
    """WeightedPartialCE's nn.CrossEntropyLoss (manual=False, C == 2) and nn.BCEWithLogitsLoss (num_classes == 1) branches."""

    @staticmethod
    def forward(ctx, mode: int, full: bool, yhat: Tensor, ys: Tensor):
        B, Cc, H, W = yhat.shape
        scratch = torch.empty(2, dtype=torch.float64, device=yhat.device)
        out = torch.empty(1, dtype=torch.float32, device=yhat.device)
        _lib.check("octave_wpce_alt_fwd", _lib.lib.octave_wpce_alt_fwd(mode, yhat.data_ptr(), ys.data_ptr(), B, Cc, H, W, int(full),
                                                                       scratch.data_ptr(), out.data_ptr(), _stream_ptr()))
        ctx.mode, ctx.full = mode, full
        ctx.save_for_backward(yhat, ys)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g_out):
        yhat, ys = ctx.saved_tensors
        B, Cc, H, W = yhat.shape
        g = torch.empty_like(yhat)
        gs = g_out.reshape(1).float().contiguous()
        _lib.check("octave_wpce_alt_bwd", _lib.lib.octave_wpce_alt_bwd(ctx.mode, yhat.data_ptr(), ys.data_ptr(), B, Cc, H, W, int(ctx.full),
                                                                       gs.data_ptr(), g.data_ptr(), _stream_ptr()))
        return None, None, g, None


class WeightedPartialCE(nn.Module):
    """Weighted partial cross-entropy on scribbles (reference: segmentor/losses.py:11-61).

    `y_hat` holds probabilities (the reference applies log(y_hat + 1e-12) directly, losses.py:52);
    a pixel is unlabelled when its one-hot row in `ys` is all zero.  The `manual=True`, `num_classes > 1` branch is the one
    on OctaScribbleNet's path (models/octa.py:52) and runs in the fused K9 kernel; the constructor default `manual=False`
    (nn.CrossEntropyLoss on ys[:,1:], losses.py:40-45,58 — well-formed for two classes only, as in the reference) and
    `num_classes == 1` (nn.BCEWithLogitsLoss, :48-49) have their own small kernels.
    """

    def __init__(self, num_classes, eps=1e-12, manual: bool = False):
        super().__init__()
        self.num_classes = num_classes
        self.eps = eps
        self.manual = manual

    def forward(self, y_hat: Tensor, ys: Tensor, ignore_bg: bool = False, reduction: str = 'mean', **kwargs) -> Tensor:
        assert y_hat.shape[1] == ys.shape[1], 'Number of class mismatch.'
        if reduction not in ('mean', 'sum'):
            raise ValueError(f'Unknown reduction {reduction}')
        if ignore_bg:
            ys[:, 0] = 0  # in-place on the caller's tensor, as the reference does (losses.py:29-30)
        if self.num_classes == 1 or not self.manual:
            _require_cuda(y_hat, "y_hat"); _require_cuda(ys, "ys")
            if self.num_classes == 1:
                if not self.manual:
                    # ys[:, 1:] of a one-channel target is empty: the reference fails inside BCEWithLogitsLoss (losses.py:40,49)
                    raise ValueError("Target size (torch.Size([0])) must be the same as input size "
                                     f"(torch.Size([{y_hat.numel()}, 1]))")
                mode = 1
            else:
                if y_hat.shape[1] != 2:
                    # `(b h w c)` targets have (C-1) entries per pixel: nn.CrossEntropyLoss rejects them unless C == 2 (losses.py:44,58)
                    raise ValueError(f"Expected input batch_size ({y_hat.numel() // y_hat.shape[1]}) to match target batch_size "
                                     f"({ys[:, 1:].numel()}).")
                mode = 0
            return _WpceAltFn.apply(mode, bool(kwargs.get('full', False)), y_hat.contiguous().float(), ys.contiguous().float())
        flags = _lib.LOSS_WPCE
        if kwargs.get('full', False):
            flags |= _lib.LOSS_WPCE_FULL
        if kwargs.get('from_logits', False):  # extension: fuse the caller's softmax(dim=1)
            flags |= _lib.LOSS_FROM_LOGITS
        out = fused_loss(_LossCfg(flags, wpce_scale_mode=reduction), yhat=y_hat, ys=ys)
        return out[_OUT_WPCE]


class DiceLoss(nn.Module):
    """Soft Dice, per sample then batch mean (reference: segmentor/losses.py:64-74)."""

    def __init__(self, eps: float = 1e-12):
        super().__init__()
        self.eps = eps

    def forward(self, input: Tensor, target: Tensor):
        out = fused_loss(_LossCfg(_lib.LOSS_DICE, dice_eps=self.eps), yhat=input, ys=target)
        return out[_OUT_DICE]


class InterlayerDivergence(nn.Module):
    """KL divergence between the full-resolution attention and the coarser ones
    (reference: segmentor/losses.py:90-172; KLD / mode='mean' branch :128-147)."""

    def __init__(self, mode='mean', eps: float = 1e-12, upscaling_mode='nn', stop_gradient: bool = False,
                 divergence='KLD'):
        super().__init__()
        assert mode in ['mean', 'sum'], f'mode {mode} is not exists/implemented.'
        self.mode = mode
        self.eps = eps
        self.stop_gradient = stop_gradient
        self.divergence = divergence

    def forward(self, attentions: Sequence[Tensor], weights: Optional[list] = None) -> Tensor:
        n_post = len(attentions) - 1
        if weights is None:
            weights = [1 for _ in range(n_post)]
        elif len(weights) != n_post:
            weights = weights[:len(attentions)]  # reference truncation quirk (losses.py:121-123)
        if self.divergence == 'KLD':
            if self.mode == 'sum':
                raise NotImplementedError('Not implemented yet.')
        elif self.divergence == 'JSD':
            pass                                   # Jensen-Shannon branch (losses.py:154-169): generic fp32 kernel
        else:
            raise NotImplementedError(f'Invalid divergence type / Not implemented: {self.divergence}')
        used = [float(w) for _, w in zip(attentions[1:], weights)]
        if len(attentions) > 5:
            raise NotImplementedError('octave_b200: at most 5 attention maps are supported')
        flags = _lib.LOSS_KLD | (_lib.LOSS_KLD_STOPGRAD if self.stop_gradient else 0)
        atts = list(attentions[:1 + len(used)])
        if self.divergence == 'JSD':
            flags |= _lib.LOSS_JSD
            atts = [a.float() for a in atts]       # the generic kernel computes on fp32 maps
        cfg = _LossCfg(flags, att_weights=used, sum_weights=float(sum(weights)), jsd_eps=float(self.eps))
        out = fused_loss(cfg, att=atts)
        if config.nan_check and bool(out[_OUT_NAN].item()):
            _log_error(f'Divergence: {out[_OUT_KLD]}')
            raise Exception('Divergence is NaN')
        return out[_OUT_KLD]


class LSDiscriminatorialLoss(nn.Module):
    """0.5*mean((y_real-1)^2) + 0.5*mean((y_fake+1)^2) (reference: discriminator/losses.py:6-14)."""

    def __init__(self):
        super().__init__()

    def forward(self, y_real: Tensor, y_fake: Tensor):
        out = fused_loss(_LossCfg(_lib.LOSS_LSD), d_real=y_real, d_fake=y_fake)
        return out[_OUT_LSD]


class LSGeneratorLoss(nn.Module):
    """0.5*mean((y_fake-1)^2) (reference: discriminator/losses.py:17-24)."""

    def __init__(self):
        super().__init__()

    def forward(self, y_fake: Tensor):
        out = fused_loss(_LossCfg(_lib.LOSS_LSG), d_fake=y_fake)
        return out[_OUT_LSG]


class FusedSegmentorLoss(nn.Module):
    """All G-step loss terms in one statistics launch + one gradient launch.

    forward(agg_map, ys, attentions, y_fake=None) -> dict of 0-d tensors
      'supervised' : WeightedPartialCE(softmax(agg_map,1), ys)   (or DiceLoss when weakly_supervise=False)
      'divergence' : InterlayerDivergence()(attentions)
      'generator'  : LSGeneratorLoss()(y_fake)                  (when y_fake is given)
    `agg_map` holds the segmentor's logits; the softmax over classes is fused into the kernel.
    """

    def __init__(self, weakly_supervise: bool = True, from_logits: bool = True, dice_eps: float = 1e-12,
                 att_weights: Optional[Sequence[float]] = None):
        super().__init__()
        self.weakly_supervise = weakly_supervise
        self.from_logits = from_logits
        self.dice_eps = dice_eps
        self.att_weights = att_weights

    def total(self, agg_map: Tensor, ys: Tensor, attentions: Sequence[Tensor], y_fake: Optional[Tensor] = None,
              lambda_sup: float = 1.0, lambda_kl: float = 1.0, lambda_g: float = 1.0):
        """The weighted G-step objective  lambda_sup * supervised + lambda_kl * divergence + lambda_g * generator  as ONE
        differentiable scalar whose gradients were written by the same pass that computed the values (K9 single pass).
        Returns the same dict as forward() (components detached) plus 'total'.  Falls back to forward() + torch
        arithmetic for configurations the single pass does not cover (Dice, non-pyramid maps)."""
        w = list(self.att_weights) if self.att_weights is not None else [1.0] * (len(attentions) - 1)
        flags = _lib.LOSS_WPCE | _lib.LOSS_KLD | (_lib.LOSS_FROM_LOGITS if self.from_logits else 0) | \
            (_lib.LOSS_LSG if y_fake is not None else 0)
        ok = self.weakly_supervise and len(attentions) >= 2
        if ok:
            dtype = _common_dtype(agg_map, *attentions)
            yhat = _prep_map(agg_map, "agg_map", dtype)
            ysm = _prep_map(ys, "ys", dtype)
            att = [_prep_map(a, f"attentions[{i}]", dtype) for i, a in enumerate(attentions)]
            fake = _prep_logits(y_fake, "y_fake") if y_fake is not None else None
            cfg = _LossCfg(flags, att_weights=w, sum_weights=float(sum(w)))
            desc = _build_desc(cfg, yhat, att, None, fake)
            ok = bool(_lib.lib.octave_loss_fused_supported(C.byref(desc)))
        if not ok:
            res = self.forward(agg_map, ys, attentions, y_fake)
            tot = lambda_sup * res['supervised'] + lambda_kl * res['divergence']
            if y_fake is not None:
                tot = tot + lambda_g * res['generator']
            res['total'] = tot
            return res
        tot, out = _FusedTotalFn.apply(cfg, (lambda_sup, lambda_kl, lambda_g), yhat, ysm, fake, *att)
        res = {'total': tot, 'supervised': out[_OUT_WPCE], 'divergence': out[_OUT_KLD], 'nan_flag': out[_OUT_NAN]}
        if y_fake is not None:
            res['generator'] = out[_OUT_LSG]
        return res

    def forward(self, agg_map: Tensor, ys: Tensor, attentions: Sequence[Tensor], y_fake: Optional[Tensor] = None):
        flags = (_lib.LOSS_WPCE if self.weakly_supervise else _lib.LOSS_DICE) | _lib.LOSS_KLD
        if self.from_logits:
            flags |= _lib.LOSS_FROM_LOGITS
        if y_fake is not None:
            flags |= _lib.LOSS_LSG
        w = list(self.att_weights) if self.att_weights is not None else [1.0] * (len(attentions) - 1)
        cfg = _LossCfg(flags, dice_eps=self.dice_eps, att_weights=w, sum_weights=float(sum(w)))
        out = fused_loss(cfg, yhat=agg_map, ys=ys, att=list(attentions), d_fake=y_fake)
        res = {'supervised': out[_OUT_WPCE if self.weakly_supervise else _OUT_DICE], 'divergence': out[_OUT_KLD],
               'nan_flag': out[_OUT_NAN]}
        if y_fake is not None:
            res['generator'] = out[_OUT_LSG]
        return res


def _log_error(msg: str) -> None:
    try:
        from loguru import logger
        logger.error(msg)
    except Exception:  # pragma: no cover
        import logging
        logging.getLogger("octave_b200").error(msg)
