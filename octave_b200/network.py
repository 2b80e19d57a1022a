"""Host-side mirror of the reference network (same module tree, parameter names, shapes and init order, so
`state_dict()` round-trips with the reference and a seeded construction yields identical weights), with the
arithmetic done by the sm_100a kernels through explicit forward / backward passes over NHWC views.

Reference modules restated here (paths under /root/reference/architectures/):
  extra/resnest.py      ResNestDecoder :18-43, Upsampling :46-54, SplAtConv2d :57-138, Bottleneck :170-267,
                        ResNet (deep stem, avg_down, avd) :277-429, resnest50 :451-459
  segmentor/blocks.py   AdversarialAttentionGate :12-46
  segmentor/compose.py  ResnestUNet :12-199

Each block exposes  fwd(x: Act, ...) -> (y: Act, ctx)  and  bwd(ctx, dy: Act, grads: dict) -> dx: Act ;
`grads` maps nn.Parameter -> fp32 gradient.  nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d objects are used as
parameter containers only (their forward is never called).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn, Tensor
from torch.nn import BatchNorm2d, Conv2d, ConvTranspose2d, ReLU

from . import config, ops
from .ops import Act, ConvSpec

Grads = Dict[nn.Parameter, Tensor]


def compute_dtype() -> torch.dtype:
    return torch.bfloat16 if config.compute_dtype == "bf16" else torch.float32


def _acc(grads: Grads, p: Optional[nn.Parameter], g: Optional[Tensor]) -> None:
    if p is None or g is None:
        return
    g = g.reshape(p.shape)
    grads[p] = g if p not in grads else grads[p] + g


def _spec(conv: nn.Module) -> ConvSpec:
    s = getattr(conv, "_oct_spec", None)
    if s is None:
        if isinstance(conv, ConvTranspose2d):
            s = ConvSpec(conv.weight, conv.bias, conv.in_channels, conv.out_channels, 2, 2, 0, 1, transposed=True)
        else:
            s = ConvSpec(conv.weight, conv.bias, conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0],
                         conv.padding[0], conv.groups)
        conv._oct_spec = s
    return s


# ---------------------------------------------------------------------------------------------------
# BatchNorm2d (+ReLU, +residual) on an Act
# ---------------------------------------------------------------------------------------------------
def bn_fwd(bn: BatchNorm2d, z: Act, training: bool, relu: bool, res: Optional[Act] = None, out: Optional[Act] = None,
           want_gap: bool = False, sums: Optional[Tensor] = None):
    """y = act(BN(z) + res).  `sums`: per-channel statistics already produced by the conv epilogue.  -> (y, ctx, gap)"""
    if training:
        if sums is None:
            sums = ops.chan_stats(z)
        ab, mi = ops.bn_prepare(z.C, z.npix, sums, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                bn.num_batches_tracked, bn.eps, bn.momentum, True, z.device)
    else:
        ab, mi = ops.bn_prepare(z.C, z.npix, None, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                None, bn.eps, bn.momentum, False, z.device)
    y, gap = ops.affine_act(z, ab, res, relu, out, want_gap)
    return y, (bn, z, mi, ab, training, relu and res is None, y), gap


def bn_bwd(ctx, dy: Act, mask: Optional[Act], grads: Grads, out: Optional[Act] = None, dmasked: Optional[Act] = None) -> Act:
    """mask: the activation whose ReLU gates dy.  When it is this BN's own output (y = relu(BN(z)), no residual) the
    mask is recomputed from z inside the kernels instead of being read (one tensor pass less in each kernel)."""
    bn, z, mi, ab, training, relu_self, y = ctx
    relu_ab = None
    if mask is not None and relu_self and mask is y:
        mask, relu_ab = None, ab
    dz, dg, db = ops.bn_bwd(dy, mask, z, mi, bn.weight.detach(), training, out, relu_ab, dmasked)
    _acc(grads, bn.weight, dg)
    _acc(grads, bn.bias, db)
    return dz


# ---------------------------------------------------------------------------------------------------
# Inference: conv -> BatchNorm2d(eval) [-> ReLU] as ONE convolution (reference: ResnestUNet.predict, compose.py:189-199)
# ---------------------------------------------------------------------------------------------------
_fold_active = False     # set by ResnestUNet.forward for an eval-mode pass that records no tape


def _folded_spec(conv: Conv2d, bn: BatchNorm2d) -> ConvSpec:
    """ConvSpec of  W' = W * g/sqrt(var+eps) (per output channel),  b' = beta + (b - mean) * g/sqrt(var+eps):
    eval-mode BN(conv(x)) == conv'(x).  Cached until a parameter or running statistic changes."""
    key = (conv.weight._version, conv.weight.data_ptr(), bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, None if conv.bias is None else conv.bias._version)
    hit = getattr(conv, "_oct_fold", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        a = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        b = bn.bias.detach().float() - bn.running_mean.float() * a
        if conv.bias is not None:
            b = b + conv.bias.detach().float() * a
        w = (conv.weight.detach().float() * a.view(-1, 1, 1, 1)).contiguous()
    spec = ConvSpec(w, b.contiguous(), conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0],
                    conv.padding[0], conv.groups)
    conv._oct_fold = (key, spec)
    return spec


def conv_bn_fwd(conv: Conv2d, bn: BatchNorm2d, x: Act, training: bool, relu: bool, out: Optional[Act] = None):
    """y = act(BN(conv(x))) -> (y, bn ctx).  Under `_fold_active` (eval, no tape) the BN scale lives in the weights and
    the shift + ReLU in the conv epilogue: one kernel and no pre-activation tensor; ctx is None."""
    if _fold_active and not training:
        return ops.conv_fwd(x, _folded_spec(conv, bn), out=out, act=1 if relu else 0), None
    z, s = ops.conv_fwd(x, _spec(conv), want_stats=training)
    y, c, _ = bn_fwd(bn, z, training, relu, out=out, sums=s)
    return y, c


_side_streams: Dict[torch.device, "torch.cuda.Stream"] = {}
_side_busy = False


def on_side_stream(fn, *acts: Act):
    """Run fn() on the side stream after the main stream's work so far; `acts` are the buffers it reads (marked as used by
    that stream, so the caching allocator does not hand their memory out again before the kernels have run)."""
    global _side_busy
    dev = (acts[0].buf if isinstance(acts[0], Act) else acts[0]).device if acts else torch.device("cuda", torch.cuda.current_device())
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = fn()
    for a in acts:
        (a.buf if isinstance(a, Act) else a).record_stream(side)
    _side_busy = True
    return out


def _overlap(x: Act) -> bool:
    return bool(config.overlap_wgrad and x.buf.is_cuda and x.dtype == torch.bfloat16)


def _wgrad(x: Act, dz: Act, spec: ConvSpec, zero_bias_grad: bool):
    """Weight gradient of one conv, on the side stream when config.overlap_wgrad (see config.py)."""
    if not (config.overlap_wgrad and x.buf.is_cuda and spec.tc_ok(x.dtype)):
        return ops.conv_wgrad(x, dz, spec, zero_bias_grad)
    return on_side_stream(lambda: ops.conv_wgrad(x, dz, spec, zero_bias_grad), x, dz)


def join_side_stream(grads: Optional[Grads] = None) -> None:
    """The main stream waits for every weight gradient issued on the side stream; gradients allocated there are handed to
    the main stream's allocator bookkeeping."""
    global _side_busy
    if not _side_busy:
        return
    main = torch.cuda.current_stream()
    side = _side_streams.get(torch.device("cuda", main.device_index))
    if side is not None:
        main.wait_stream(side)
        if grads is not None:
            for g in grads.values():
                if g is not None and g.is_cuda:
                    g.record_stream(main)
    _side_busy = False


def conv_bwd(conv: nn.Module, x: Act, dz: Act, grads: Grads, need_dx: bool = True, dx_out: Optional[Act] = None,
             accumulate: bool = False, zero_bias_grad: bool = False) -> Optional[Act]:
    """zero_bias_grad: the conv feeds a training-mode BatchNorm, whose input gradient sums to zero over every channel,
    so the bias gradient is identically zero (the reference computes rounding noise there) and its reduction is skipped."""
    spec = _spec(conv)
    dw, db = _wgrad(x, dz, spec, zero_bias_grad)
    _acc(grads, conv.weight, dw)
    _acc(grads, conv.bias, db)
    if not need_dx:
        return None
    return ops.conv_dgrad(dz, spec, x.H, x.W, dx_out, accumulate)


# ---------------------------------------------------------------------------------------------------
class SplAtConv2d(nn.Module):
    """Split-attention conv, radix 2 (reference: extra/resnest.py:57-138)."""

    def __init__(self, in_channels, channels, kernel_size, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1,
                 bias=True, radix=2, reduction_factor=4, norm_layer=None, **kwargs):
        super().__init__()
        assert radix == 2, "octave_b200 ports radix=2 (resnest50 / ResNestDecoder configuration)"
        inter_channels = max(in_channels * radix // reduction_factor, 32)
        self.radix, self.cardinality, self.channels = radix, groups, channels
        self.conv = Conv2d(in_channels, channels * radix, kernel_size, stride, padding, dilation, groups=groups * radix,
                           bias=bias, **kwargs)
        self.use_bn = norm_layer is not None
        self.bn0 = norm_layer(channels * radix)
        self.relu = ReLU(inplace=True)
        self.fc1 = Conv2d(channels, inter_channels, 1, groups=self.cardinality)
        self.bn1 = norm_layer(inter_channels)
        self.fc2 = Conv2d(inter_channels, channels * radix, 1, groups=self.cardinality)

    def fwd(self, x: Act, relu_out: bool, out: Optional[Act] = None):
        tr = self.training
        z, zs = ops.conv_fwd(x, _spec(self.conv), want_stats=tr)                # resnest.py:99
        U, bn0ctx, gap = bn_fwd(self.bn0, z, tr, True, want_gap=True, sums=zs)  # :101-116 (radix sum + GAP fused)
        hw = float(x.H * x.W)
        card = self.cardinality
        w1 = self.fc1.weight.detach().reshape(self.fc1.out_channels, -1)
        w2 = self.fc2.weight.detach().reshape(self.fc2.out_channels, -1)
        bn1 = self.bn1
        if ops.attn_fused_ok(gap.shape[0], gap.shape[1], w1.shape[0], card, self.radix) and w2.shape[0] == 2 * gap.shape[1]:
            # two launches: fc1 + bn1 + relu, fc2 + r-softmax (the whole batch column of an output sits in one warp)
            h1, h1n, mi1 = ops.glinear_bn_relu_fwd(gap, w1, self.fc1.bias.detach(), 1.0 / hw, bn1.weight.detach(), bn1.bias.detach(),
                                                   bn1.running_mean, bn1.running_var, bn1.num_batches_tracked if tr else None,
                                                   bn1.eps, bn1.momentum, tr)             # :118-122
            att = ops.glinear_rsoftmax_fwd(h1n, w2, self.fc2.bias.detach(), gap.shape[1])   # :125-127
        else:
            h1 = ops.glinear_fwd(gap, w1, self.fc1.bias.detach(), card, 1.0 / hw)   # :118
            h1n, mi1 = ops.bn1d_relu_fwd(h1, bn1.weight.detach(), bn1.bias.detach(), bn1.running_mean, bn1.running_var,
                                         bn1.num_batches_tracked if tr else None, bn1.eps, bn1.momentum, tr)  # :120-122
            logits = ops.glinear_fwd(h1n, w2, self.fc2.bias.detach(), card, 1.0)    # :125
            att = ops.rsoftmax_fwd(logits, self.radix)                              # :125-127 view(B,radix,C) softmax(dim=1)
        o = ops.splat_combine(U, att, relu_out, out)                            # :133-135
        return o, (x, z, U, bn0ctx, gap, h1, h1n, mi1, att, o if relu_out else None, tr, hw)

    def bwd(self, ctx, dout: Act, grads: Grads, need_dx: bool = True) -> Optional[Act]:
        x, z, U, bn0ctx, gap, h1, h1n, mi1, att, mask, tr, hw = ctx
        card = self.cardinality
        datt = ops.splat_bwd_reduce(dout, mask, U)
        w2 = self.fc2.weight.detach().reshape(self.fc2.out_channels, -1)
        side = on_side_stream if _overlap(x) else None
        if (config.fuse_attention_branch_bwd and ops.attn_fused_ok(gap.shape[0], gap.shape[1], h1.shape[1], card, self.radix)
                and w2.shape[0] == 2 * gap.shape[1]):
            dh1, dg1, dbt1, dw2, db2 = ops.attn_bwd_fused(datt, att, w2, h1, h1n, self.bn1.weight.detach(), mi1, tr, card, side=side)
        else:
            dlogits = ops.rsoftmax_bwd(datt, att, self.radix)
            dh1n, dw2, db2 = ops.glinear_bwd(dlogits, h1n, w2, card, 1.0, side=side)
            dh1, dg1, dbt1 = ops.bn1d_relu_bwd(dh1n, h1, h1n, self.bn1.weight.detach(), mi1, tr)
        _acc(grads, self.fc2.weight, dw2); _acc(grads, self.fc2.bias, db2)
        _acc(grads, self.bn1.weight, dg1); _acc(grads, self.bn1.bias, dbt1)
        w1 = self.fc1.weight.detach().reshape(self.fc1.out_channels, -1)
        dgap, dw1, db1 = ops.glinear_bwd(dh1, gap, w1, card, 1.0 / hw, side=side)
        _acc(grads, self.fc1.weight, dw1); _acc(grads, self.fc1.bias, db1)
        bn0, z0, mi0, ab0, tr0, _relu_self, _u = bn0ctx
        if dout.C // 8 <= 256:
            # dU is never materialised: both BatchNorm-backward passes rebuild it from dout while streaming z
            dz, dg0, db0 = ops.splat_bn_bwd(dout, mask, att, dgap, 1.0, z0, ab0, mi0, bn0.weight.detach(), tr0)
            _acc(grads, bn0.weight, dg0); _acc(grads, bn0.bias, db0)
        else:
            dU = ops.splat_bwd_du(dout, mask, att, dgap, 1.0)
            dz = bn_bwd(bn0ctx, dU, U, grads, out=dU)
        return conv_bwd(self.conv, x, dz, grads, need_dx, zero_bias_grad=tr)


class Bottleneck(nn.Module):
    """ResNeSt bottleneck (reference: extra/resnest.py:170-267), radix 2, avd after the split-attention conv."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, radix=2, cardinality=1, bottleneck_width=64,
                 avd=False, avd_first=False, is_first=False, norm_layer=None):
        super().__init__()
        group_width = int(planes * (bottleneck_width / 64.)) * cardinality
        self.conv1 = Conv2d(inplanes, group_width, kernel_size=1, bias=False)
        self.bn1 = norm_layer(group_width)
        self.radix = radix
        self.avd = avd and (stride > 1 or is_first)
        assert not avd_first
        if self.avd:
            self.avd_layer = nn.AvgPool2d(3, stride, padding=1)
            self._avd_pd = ops.pool_desc("avg", 3, stride, 1, False, True)
            stride = 1
        self.conv2 = SplAtConv2d(group_width, group_width, kernel_size=3, stride=stride, padding=1, dilation=1,
                                 groups=cardinality, bias=False, radix=radix, norm_layer=norm_layer)
        self.conv3 = Conv2d(group_width, planes * 4, kernel_size=1, bias=False)
        self.bn3 = norm_layer(planes * 4)
        self.relu = ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def _downsample_fwd(self, x: Act, tr: bool):
        pool, convd, bnd = self.downsample[0], self.downsample[1], self.downsample[2]
        pd = ops.pool_desc("avg", pool.kernel_size, pool.stride, 0, True, False)
        r = x
        if pool.kernel_size != 1:
            r, _ = ops.pool_fwd(pd, x)
        res, cbd = conv_bn_fwd(convd, bnd, r, tr, False)
        return res, (pd, r, cbd)

    def fwd(self, x: Act, out: Optional[Act] = None):
        tr = self.training
        cd = None
        y1, c1 = conv_bn_fwd(self.conv1, self.bn1, x, tr, True)
        s, c2 = self.conv2.fwd(y1, relu_out=False)
        if self.avd:
            sp, _ = ops.pool_fwd(self._avd_pd, s)
        else:
            sp = s
        z3, s3 = ops.conv_fwd(sp, _spec(self.conv3), want_stats=tr)
        if self.downsample is not None:
            # (running this shortcut on the side stream beside the main branch was measured: no gain, 45.27 vs 45.17 ms/step)
            res, cd = self._downsample_fwd(x, tr)
        else:
            res = x
        y, c3, _ = bn_fwd(self.bn3, z3, tr, True, res=res, out=out, sums=s3)
        return y, (x, y1, c1, c2, s, sp, c3, cd, y)

    def bwd(self, ctx, dy: Act, grads: Grads) -> Act:
        x, y1, c1, c2, s, sp, c3, cd, y = ctx
        # identity shortcut: its gradient dy * (y > 0) comes out of the same pass as bn3's input gradient
        dres = dy.like() if cd is None else None
        dz3 = bn_bwd(c3, dy, y, grads, dmasked=dres)
        dsp = conv_bwd(self.conv3, sp, dz3, grads)
        ds = ops.pool_bwd(self._avd_pd, dsp, None, s.H, s.W) if self.avd else dsp
        dy1 = self.conv2.bwd(c2, ds, grads)
        dz1 = bn_bwd(c1, dy1, y1, grads, out=dy1)
        if cd is not None:
            pd, r, cbd = cd
            convd = self.downsample[1]
            dzr = bn_bwd(cbd, dy, y, grads)
            dr = conv_bwd(convd, r, dzr, grads)
            dx = ops.pool_bwd(pd, dr, None, x.H, x.W) if r is not x else dr
        else:
            dx = dres
        conv_bwd(self.conv1, x, dz1, grads, dx_out=dx, accumulate=True)
        return dx


class ResNet(nn.Module):
    """Encoder factory mirroring reference extra/resnest.py:277-429 for the resnest50 arguments
    (radix 2, groups 1, deep stem 32, avg_down, avd).  Only construction lives here; ResnestUNet re-parents the stages."""

    def __init__(self, block, layers, radix=2, groups=1, bottleneck_width=64, num_classes=1000, deep_stem=True,
                 stem_width=32, avg_down=True, avd=True, avd_first=False, norm_layer=BatchNorm2d):
        self.cardinality, self.bottleneck_width = groups, bottleneck_width
        self.inplanes = stem_width * 2
        self.avg_down, self.radix, self.avd, self.avd_first = avg_down, radix, avd, avd_first
        super().__init__()
        self.conv1 = nn.Sequential(
            Conv2d(3, stem_width, kernel_size=3, stride=2, padding=1, bias=False),
            norm_layer(stem_width), ReLU(inplace=True),
            Conv2d(stem_width, stem_width, kernel_size=3, stride=1, padding=1, bias=False),
            norm_layer(stem_width), ReLU(inplace=True),
            Conv2d(stem_width, stem_width * 2, kernel_size=3, stride=1, padding=1, bias=False))
        self.bn1 = norm_layer(self.inplanes)
        self.relu = ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], norm_layer=norm_layer, is_first=False)
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, norm_layer=norm_layer)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, norm_layer=norm_layer)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, norm_layer=norm_layer)
        # the reference also builds (and then drops) the ImageNet classifier; it consumes the RNG, so do we
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():                                                  # resnest.py:368-374
            if isinstance(m, Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, norm_layer):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, norm_layer=None, is_first=True):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.AvgPool2d(kernel_size=stride, stride=stride, ceil_mode=True, count_include_pad=False),
                Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=1, bias=False),
                norm_layer(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample=downsample, radix=self.radix,
                        cardinality=self.cardinality, bottleneck_width=self.bottleneck_width, avd=self.avd,
                        avd_first=self.avd_first, is_first=is_first, norm_layer=norm_layer)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, radix=self.radix, cardinality=self.cardinality,
                                bottleneck_width=self.bottleneck_width, avd=self.avd, avd_first=self.avd_first,
                                norm_layer=norm_layer))
        return nn.Sequential(*layers)


def resnest50(pretrained=False, **kwargs):
    model = ResNet(Bottleneck, [3, 4, 6, 3], radix=2, groups=1, bottleneck_width=64, deep_stem=True, stem_width=32,
                   avg_down=True, avd=True, avd_first=False)
    model_path = kwargs.get('model_path', './models/resnest50-528c19ca.pth')
    if pretrained:
        model.load_state_dict(torch.load(model_path))
    return model


def layer_fwd(layer: nn.Sequential, x: Act, out: Optional[Act] = None):
    ctxs = []
    n = len(layer)
    for i, blk in enumerate(layer):
        x, c = blk.fwd(x, out if i == n - 1 else None)
        ctxs.append(c)
    return x, ctxs


def layer_bwd(layer: nn.Sequential, ctxs, dy: Act, grads: Grads) -> Act:
    for blk, c in zip(reversed(list(layer)), reversed(ctxs)):
        dy = blk.bwd(c, dy, grads)
    return dy


# ---------------------------------------------------------------------------------------------------
class ResNestDecoder(nn.Module):
    """Decoder block (reference: extra/resnest.py:18-43)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Sequential(
            Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False),
            BatchNorm2d(out_channels),
            ReLU(inplace=True),
            SplAtConv2d(out_channels, out_channels, kernel_size=3, padding=1, stride=1, groups=2, radix=2,
                        norm_layer=BatchNorm2d),
            ReLU(inplace=True))
        self.downsample = nn.Sequential(Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=False),
                                        BatchNorm2d(out_channels))
        self.relu = ReLU(inplace=True)

    def fwd(self, x: Act, out: Optional[Act] = None):
        tr = self.training
        zr, sr = ops.conv_fwd(x, _spec(self.downsample[0]), want_stats=tr)
        y0, c0 = conv_bn_fwd(self.conv[0], self.conv[1], x, tr, True)
        s, cs = self.conv[3].fwd(y0, relu_out=True)
        y, cr, _ = bn_fwd(self.downsample[1], zr, tr, True, res=s, out=out, sums=sr)   # relu(BN(shortcut) + relu(splat))
        return y, (x, y0, c0, cs, s, cr, y)

    def bwd(self, ctx, dy: Act, grads: Grads) -> Act:
        x, y0, c0, cs, s, cr, y = ctx
        dsum = dy.like()
        dzr = bn_bwd(cr, dy, y, grads, dmasked=dsum)      # dsum = dy * (y > 0): gradient of the split-attention branch
        dy0 = self.conv[3].bwd(cs, dsum, grads)
        dz0 = bn_bwd(c0, dy0, y0, grads, out=dy0)
        dx = conv_bwd(self.conv[0], x, dz0, grads)
        conv_bwd(self.downsample[0], x, dzr, grads, dx_out=dx, accumulate=True)
        return dx


class Upsampling(nn.Module):
    """ConvTranspose2d k2 s2 (reference: extra/resnest.py:46-54); writes straight into a concat slice."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.up = ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2)

    def fwd(self, x: Act, out: Optional[Act] = None):
        y = ops.conv_fwd(x, _spec(self.up), out)
        return y, (x,)

    def bwd(self, ctx, dy: Act, grads: Grads) -> Act:
        (x,) = ctx
        side = on_side_stream if (config.overlap_wgrad and x.buf.is_cuda and _spec(self.up).tc_ok(x.dtype)) else None
        dx, dw, db = ops.convt_bwd(x, dy, _spec(self.up), side=side)
        _acc(grads, self.up.weight, dw); _acc(grads, self.up.bias, db)
        return dx


class AdversarialAttentionGate(nn.Module):
    """1x1 conv -> softmax over classes -> x * sum_{c>=1} p_c (reference: segmentor/blocks.py:12-46)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv1 = Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=(1, 1), stride=1)
        self.softmax = nn.Softmax(dim=1)

    def _w(self):
        return self.conv1.weight.detach().reshape(self.conv1.out_channels, -1).float().contiguous(), self.conv1.bias.detach().float().contiguous()

    def fwd(self, x: Act):
        w, b = self._w()
        y_hat, gated = ops.head_fwd(x, w, b, 1)
        return gated, y_hat, (x, w, b)

    def bwd(self, ctx, dgated: Act, dyhat: Optional[Tensor], grads: Grads) -> Act:
        x, w, b = ctx
        dx, dw, db = ops.head_bwd(x, w, b, 1, dyhat, dgated)
        _acc(grads, self.conv1.weight, dw); _acc(grads, self.conv1.bias, db)
        return dx


class GlobalAveragePooling2D(nn.Module):
    """Parameter-free pooling used only by the (unported, inert) classification heads."""

    def forward(self, x: Tensor):
        return x.mean(dim=(2, 3))


# ---------------------------------------------------------------------------------------------------
class _SegmentorFn(torch.autograd.Function):
    """Whole-network autograd node: one explicit forward pass and one explicit backward pass over the kernels."""

    @staticmethod
    def forward(ctx, net: "ResnestUNet", x: Tensor, *params):
        outs, tape = net._fwd(x)
        ctx.net, ctx.tape, ctx.params = net, tape, params
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        net = ctx.net
        if ctx.tape is None:
            raise RuntimeError("octave_b200: backward through the segmentor twice (tape already released)")
        grads: Grads = {}
        net._bwd(ctx.tape, gouts, grads)
        join_side_stream(grads)
        ctx.tape = None
        if net._grad_ready_hook is not None:
            # Every gradient has been handed to the hook (the data-parallel reducer), whose owner assigns param.grad when the
            # exchange is done (GradAllReducer.finish).  Returning the tensors as well would make autograd's AccumulateGrad
            # CLONE each of them (the reducer still holds references, so it cannot steal them): ~400 copy kernels per step.
            return (None, None, *[None for _ in ctx.params])
        return (None, None, *[grads.get(p) for p in ctx.params])


class ResnestUNet(nn.Module):
    """ResNeSt-50 encoder + 5-level decoder U-Net with adversarial attention gates
    (reference: segmentor/compose.py:12-199).  Returns (attentions full-res first, agg_map logits, x_4)."""

    def __init__(self, num_classes: int, pretrain: bool, weight_path: str = None, gating_level: int = 4,
                 encoder_gating: bool = False):
        super().__init__()
        if encoder_gating:
            raise NotImplementedError("octave_b200: encoder_gating=True (compose.py:28-37) is outside the ported path "
                                      "(OctaScribbleNet default is False, models/octa.py:27)")
        resnest = resnest50(pretrained=pretrain, model_path=weight_path)
        self.gating_level = gating_level
        self.encoder_gating = encoder_gating
        self.num_classes = num_classes
        self.encoder_0_1_2 = nn.Sequential(resnest.conv1, resnest.bn1, resnest.relu)
        self.encoder_0_2_2 = resnest.maxpool
        self.upsampling_0 = Upsampling(64, 64)
        self.decoder_0 = ResNestDecoder(64, 32)
        self.aag_0 = AdversarialAttentionGate(32, num_classes)
        self.encoder_1 = resnest.layer1
        self.upsampling_1 = Upsampling(256, 64)
        self.decoder_1 = ResNestDecoder(128, 64)
        self.aag_1 = AdversarialAttentionGate(64, num_classes)
        self.encoder_2 = resnest.layer2
        self.aag_2 = AdversarialAttentionGate(256, num_classes)
        self.upsampling_2 = Upsampling(512, 256)
        self.decoder_2 = ResNestDecoder(512, 256)
        self.encoder_3 = resnest.layer3
        self.upsampling_3 = Upsampling(1024, 512)
        self.decoder_3 = ResNestDecoder(1024, 512)
        self.aag_3 = AdversarialAttentionGate(512, num_classes)
        self.encoder_4 = resnest.layer4
        self.upsampling_4 = Upsampling(2048, 1024)
        self.decoder_4 = ResNestDecoder(2048, 1024)
        self.aag_4 = AdversarialAttentionGate(1024, num_classes)
        self.fc = Conv2d(in_channels=32, out_channels=num_classes, kernel_size=1, stride=1)
        # Inert classification heads: kept so that state_dict keys match the reference.
        self.linear_head_emb = nn.Sequential(GlobalAveragePooling2D(), nn.Linear(2048, num_classes))
        self.linear_head_dec = nn.Sequential(
            nn.AdaptiveAvgPool2d((32, 32)), Conv2d(in_channels=num_classes, out_channels=64, kernel_size=7),
            ReLU(inplace=True), BatchNorm2d(num_features=64), Conv2d(in_channels=64, out_channels=512, kernel_size=7),
            ReLU(inplace=True), BatchNorm2d(num_features=512), GlobalAveragePooling2D(), nn.Linear(512, num_classes))
        self._maxpool_pd = ops.pool_desc("max", 3, 2, 1, False, True)
        self._grad_ready_hook = None

    # ---- explicit passes ---------------------------------------------------------------------------
    def _hot_params(self) -> List[nn.Parameter]:
        return [p for n, p in self.named_parameters() if not n.startswith("linear_head_")]

    def _stem_fwd(self, x, out: Act):
        """x: Act (fp32 mode, NHWC with 3 of 8 channels) or, in bf16 mode, the raw NCHW fp32 tensor: the 3x3 stride-2
        conv then runs on the tensor cores as a 3x3 stride-1 conv over the space-to-depth input (see disc.cu)."""
        seq, bn1 = self.encoder_0_1_2[0], self.encoder_0_1_2[1]
        tr = self.training
        s0 = None
        if isinstance(x, Act):
            z0 = ops.conv_fwd(x, _spec(seq[0]))
        else:
            B, _, H, W = x.shape
            xs = Act.empty(B, H // 2, W // 2, 32, torch.bfloat16, x.device)      # nchw_to_s2d writes every channel (pads as zeros)
            ops.nchw_to_s2d(x, xs, 8, 0)
            s0 = ops.zeros_f64(2 * seq[0].out_channels, x.device) if tr else None
            z0 = ops.conv4x4s2_tc_fwd(xs, ops.pack_weight_s2d(seq[0].weight.detach(), None, 0, 8), None, seq[0].out_channels,
                                      H // 2, W // 2, 0, stats=s0)
            x = xs
        y0, c0, _ = bn_fwd(seq[1], z0, tr, True, sums=s0)
        y1, c1 = conv_bn_fwd(seq[3], seq[4], y0, tr, True)
        y2, c2 = conv_bn_fwd(seq[6], bn1, y1, tr, True, out=out)
        return y2, (x, y0, c0, y1, c1, c2, y2)

    def _stem_bwd(self, ctx, dy: Act, grads: Grads) -> None:
        seq = self.encoder_0_1_2[0]
        x, y0, c0, y1, c1, c2, y2 = ctx
        dz2 = bn_bwd(c2, dy, y2, grads)
        dy1 = conv_bwd(seq[6], y1, dz2, grads)
        dz1 = bn_bwd(c1, dy1, y1, grads, out=dy1)
        dy0 = conv_bwd(seq[3], y0, dz1, grads)
        dz0 = bn_bwd(c0, dy0, y0, grads, out=dy0)
        if x.C == 32:   # space-to-depth tensor-core path
            _acc(grads, seq[0].weight, ops.conv4x4s2_tc_wgrad(x, dz0, 3, 8, 3))
        else:
            conv_bwd(seq[0], x, dz0, grads, need_dx=False)

    def _fwd(self, x: Tensor):
        """x: [B,3,H,W] (any float dtype) -> ((att..., agg_map, x_4) as fp32 NCHW tensors, tape)"""
        if not x.is_cuda:
            raise RuntimeError("octave_b200: input is on CPU; the B200 kernels have no CPU fallback")
        B, Cin, H, W = x.shape
        if Cin != 3:
            raise ValueError("ResnestUNet expects 3 input channels (resnest.py:327)")
        if H % 16 or W % 16:
            raise ValueError(f"input extent {H}x{W} must be a multiple of 16 (the reference's skip concatenations "
                             f"fail otherwise, compose.py:141-169)")
        dt, dev = compute_dtype(), x.device
        tape = {}
        if dt == torch.bfloat16:
            self._repack()
        xa = ops.nchw_to_nhwc(x, dt) if dt != torch.bfloat16 else x.detach()
        cat1 = Act.empty(B, H // 2, W // 2, 128, dt, dev)
        cat2 = Act.empty(B, H // 4, W // 4, 512, dt, dev)
        cat3 = Act.empty(B, H // 8, W // 8, 1024, dt, dev)
        cat4 = Act.empty(B, H // 16, W // 16, 2048, dt, dev)
        x_0_0, tape["stem"] = self._stem_fwd(xa, cat1.slice(0, 64))               # compose.py:102
        x_0_1, tape["maxpool"] = ops.pool_fwd(self._maxpool_pd, x_0_0)           # :103
        x_1, tape["enc1"] = layer_fwd(self.encoder_1, x_0_1, cat2.slice(0, 256))  # :109
        x_2, tape["enc2"] = layer_fwd(self.encoder_2, x_1, cat3.slice(0, 512))
        x_3, tape["enc3"] = layer_fwd(self.encoder_3, x_2, cat4.slice(0, 1024))
        h3, w3 = x_3.H, x_3.W
        if (h3 % 2) or (w3 % 2):                                                  # :125-130 zero-pad bottom/right
            x_3p = Act.empty(B, h3 + h3 % 2, w3 + w3 % 2, 1024, dt, dev)
            ops.copy_window(x_3, x_3p, False)
        else:
            x_3p = x_3
        tape["x3"] = (x_3, x_3p)
        x_4, tape["enc4"] = layer_fwd(self.encoder_4, x_3p)                       # :132
        atts = []
        # level 4: the ConvT output is cropped back to x_3's extent while being written (compose.py:140-147)
        _, tape["up4"] = self.upsampling_4.fwd(x_4, cat4.slice(1024, 1024))
        d, tape["dec4"] = self.decoder_4.fwd(cat4)
        tape["aag4"] = None
        if self.gating_level >= 4:
            d, y, tape["aag4"] = self.aag_4.fwd(d); atts.append(y)
        _, tape["up3"] = self.upsampling_3.fwd(d, cat3.slice(512, 512))
        d, tape["dec3"] = self.decoder_3.fwd(cat3)
        tape["aag3"] = None
        if self.gating_level >= 3:
            d, y, tape["aag3"] = self.aag_3.fwd(d); atts.append(y)
        _, tape["up2"] = self.upsampling_2.fwd(d, cat2.slice(256, 256))
        d, tape["dec2"] = self.decoder_2.fwd(cat2)
        tape["aag2"] = None
        if self.gating_level >= 2:
            d, y, tape["aag2"] = self.aag_2.fwd(d); atts.append(y)
        _, tape["up1"] = self.upsampling_1.fwd(d, cat1.slice(64, 64))
        d, tape["dec1"] = self.decoder_1.fwd(cat1)
        tape["aag1"] = None
        if self.gating_level >= 1:
            d, y, tape["aag1"] = self.aag_1.fwd(d); atts.append(y)
        up0, tape["up0"] = self.upsampling_0.fwd(d)                               # :175 no skip
        d, tape["dec0"] = self.decoder_0.fwd(up0)
        tape["aag0"] = None
        if self.gating_level >= 0:
            d, y, tape["aag0"] = self.aag_0.fwd(d); atts.append(y)
        wfc = self.fc.weight.detach().reshape(self.num_classes, -1).float().contiguous()
        bfc = self.fc.bias.detach().float().contiguous()
        agg_map, _ = ops.head_fwd(d, wfc, bfc, 0)                                 # :181
        tape["fc"] = (d, wfc, bfc)
        atts.reverse()                                                            # :183
        tape["n_att"] = len(atts)
        tape["dims"] = (B, H, W)
        x4_out = ops.nhwc_to_nchw(x_4)
        tape["x4"] = x_4
        return (*atts, agg_map, x4_out), tape

    def _bwd(self, tape, gouts, grads: Grads) -> None:
        n_att = tape["n_att"]
        g_att = list(gouts[:n_att])          # full-res first
        g_agg, g_x4 = gouts[n_att], gouts[n_att + 1]
        g_att.reverse()                      # now coarse (level 4) first, like the forward order
        # level index of each attention in forward order
        levels = [l for l in (4, 3, 2, 1, 0) if self.gating_level >= l]
        gy = dict(zip(levels, g_att))

        def cont(t):
            return None if t is None else t.contiguous().float()

        emitted = set()

        def emit():
            # announce the gradients completed since the last call (data-parallel buckets start their all-reduce now)
            hook = self._grad_ready_hook
            if hook is None:
                return
            # (weight gradients may still be in flight on the side stream: the reducer's stream waits for it itself, see
            # train.GradAllReducer._flush — the main stream is not held up here)
            new = [(p, g) for p, g in grads.items() if p not in emitted]
            emitted.update(grads.keys())
            if new:
                hook([p for p, _ in new], [g for _, g in new])

        d_f, wfc, bfc = tape["fc"]
        dd, dw, db = ops.head_bwd(d_f, wfc, bfc, 0, cont(g_agg), None)
        _acc(grads, self.fc.weight, dw); _acc(grads, self.fc.bias, db)
        if tape["aag0"] is not None:
            dd = self.aag_0.bwd(tape["aag0"], dd, cont(gy.get(0)), grads)
        dd = self.decoder_0.bwd(tape["dec0"], dd, grads)
        dd = self.upsampling_0.bwd(tape["up0"], dd, grads)
        emit()
        if tape["aag1"] is not None:
            dd = self.aag_1.bwd(tape["aag1"], dd, cont(gy.get(1)), grads)
        dcat1 = self.decoder_1.bwd(tape["dec1"], dd, grads)
        dd = self.upsampling_1.bwd(tape["up1"], dcat1.slice(64, 64), grads)
        emit()
        if tape["aag2"] is not None:
            dd = self.aag_2.bwd(tape["aag2"], dd, cont(gy.get(2)), grads)
        dcat2 = self.decoder_2.bwd(tape["dec2"], dd, grads)
        dd = self.upsampling_2.bwd(tape["up2"], dcat2.slice(256, 256), grads)
        emit()
        if tape["aag3"] is not None:
            dd = self.aag_3.bwd(tape["aag3"], dd, cont(gy.get(3)), grads)
        dcat3 = self.decoder_3.bwd(tape["dec3"], dd, grads)
        dd = self.upsampling_3.bwd(tape["up3"], dcat3.slice(512, 512), grads)
        emit()
        if tape["aag4"] is not None:
            dd = self.aag_4.bwd(tape["aag4"], dd, cont(gy.get(4)), grads)
        dcat4 = self.decoder_4.bwd(tape["dec4"], dd, grads)
        dx4 = self.upsampling_4.bwd(tape["up4"], dcat4.slice(1024, 1024), grads)
        emit()
        if g_x4 is not None:
            x4 = tape["x4"]
            gx = ops.nchw_to_nhwc(cont(g_x4), x4.dtype)
            ops.add_inplace(dx4, Act(gx.buf, gx.B, gx.H, gx.W, x4.C, gx.ld, 0))
        # encoder
        x_3, x_3p = tape["x3"]
        dx3p = layer_bwd(self.encoder_4, tape["enc4"], dx4, grads)
        dx3 = dcat4.slice(0, 1024)
        ops.copy_window(dx3p, dx3, True)     # crop the padding away and merge with the skip gradient
        emit()
        dx2 = layer_bwd(self.encoder_3, tape["enc3"], dx3, grads)
        ops.add_inplace(dx2, dcat3.slice(0, 512))
        emit()
        dx1 = layer_bwd(self.encoder_2, tape["enc2"], dx2, grads)
        ops.add_inplace(dx1, dcat2.slice(0, 256))
        emit()
        dx01 = layer_bwd(self.encoder_1, tape["enc1"], dx1, grads)
        B, H, W = tape["dims"]
        dx00 = ops.pool_bwd(self._maxpool_pd, dx01, tape["maxpool"], H // 2, W // 2)
        ops.add_inplace(dx00, dcat1.slice(0, 64))
        self._stem_bwd(tape["stem"], dx00, grads)
        emit()
        join_side_stream(grads)

    def _repack(self) -> None:
        """bf16 operand packs of every tensor-core conv, refreshed in one launch when the weights have changed"""
        mp = getattr(self, "_multi_packer", None)
        if mp is None or mp.specs[0].weight.device != self.fc.weight.device:
            specs = []
            for name, m in self.named_modules():
                if name.startswith("linear_head_") or m is self.fc or name.endswith((".fc1", ".fc2")):
                    continue   # heads / attention-branch linears do not run on the conv kernels
                if isinstance(m, (Conv2d, ConvTranspose2d)):
                    sp = _spec(m)
                    if sp.tc_ok(torch.bfloat16):
                        specs.append(sp)
            mp = ops.MultiPacker(specs)
            self._multi_packer = mp
        mp.run()

    # ---- public interface ----------------------------------------------------------------------------
    def forward(self, x):
        global _fold_active
        if not self.training and not torch.is_grad_enabled() and config.fold_bn_inference:
            # inference: no tape is kept, conv+BN pairs run as one convolution
            _fold_active = True
            try:
                outs, _ = self._fwd(x)
            finally:
                _fold_active = False
        else:
            outs = _SegmentorFn.apply(self, x, *self._hot_params())
        n = len(outs) - 2
        return tuple(outs[:n]), outs[n], outs[n + 1]

    def predict(self, x: Tensor, method='softmax'):
        """reference: compose.py:189-199"""
        attentions, agg_map, _ = self.forward(x)
        if method == 'softmax':
            predicate = nn.Softmax(dim=1)(agg_map)
        elif method == 'sigmoid':
            predicate = nn.Sigmoid()(agg_map)
        elif method == 'one-hot':
            predicate = torch.nn.functional.one_hot(torch.argmax(agg_map, dim=1)).permute(0, 3, 1, 2)
        elif method == 'original':
            predicate = agg_map
        return attentions, predicate

    def classification_predict(self, *a, **k):
        raise NotImplementedError("octave_b200: classification heads (compose.py:201-230) are outside the ported path")


# ---------------------------------------------------------------------------------------------------
class _BlockFn(torch.autograd.Function):
    """Autograd node around ONE block (SplAtConv2d, Bottleneck, ResNestDecoder, Upsampling, AdversarialAttentionGate):
    NCHW fp32 tensors at the boundary, kernels inside.  Used for per-module parity tests and standalone use."""

    @staticmethod
    def forward(ctx, block: nn.Module, kwargs: dict, x: Tensor, *params):
        xa = ops.nchw_to_nhwc(x, compute_dtype())
        if isinstance(block, AdversarialAttentionGate):
            gated, y_hat, c = block.fwd(xa)
            outs = (ops.nhwc_to_nchw(gated), y_hat)
        else:
            y, c = block.fwd(xa, **kwargs)
            outs = (ops.nhwc_to_nchw(y),)
        ctx.block, ctx.c, ctx.params, ctx.xa = block, c, params, xa
        ctx.set_materialize_grads(False)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *gouts):
        block, grads = ctx.block, {}
        xa = ctx.xa
        dt = compute_dtype()

        def to_act(g):
            a = ops.nchw_to_nhwc(g.contiguous().float(), dt)
            return a

        if isinstance(block, AdversarialAttentionGate):
            dg = to_act(gouts[0]) if gouts[0] is not None else Act.zeros(xa.B, xa.H, xa.W, xa.C, dt, xa.device)
            dy = gouts[1].contiguous().float() if gouts[1] is not None else None
            dx = block.bwd(ctx.c, dg, dy, grads)
        else:
            dx = block.bwd(ctx.c, to_act(gouts[0]), grads)
        gx = ops.nhwc_to_nchw(dx) if dx is not None else None
        join_side_stream(grads)
        return (None, None, gx, *[grads.get(p) for p in ctx.params])


def run_block(block: nn.Module, x: Tensor, **kwargs):
    """Apply one kernel-backed block to an NCHW tensor with autograd support."""
    params = list(block.parameters())
    return _BlockFn.apply(block, kwargs, x, *params)
