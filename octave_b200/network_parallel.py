"""The two sibling segmentors of the reference that reuse the ResnestUNet blocks with a second decoder head
(/root/reference/architectures/segmentor/compose.py): ResnestUnetParallelHead :233-361 and
ResnestUnetParallelHeadAttentionGate :364-527.  Same module tree / parameter names / construction order as the
reference (state_dict round-trips, seeded init identical); arithmetic on the sm_100a kernels through the explicit
forward / backward passes of `network.py`'s blocks.  The parallel branch (`upsampling_1_c, decoder_1_c,
upsampling_0_c, decoder_0_c, fc_c`) starts from the layer1 features x_1 and shares the stem skip x_0_0, so in the
backward pass x_1 and x_0_0 collect three gradients each."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn, Tensor
from torch.nn import Conv2d

from . import config, ops
from . import network as _n
from .network import (AdversarialAttentionGate, Grads, ResNestDecoder, ResnestUNet, Upsampling, _acc, compute_dtype,
                      layer_bwd, layer_fwd, resnest50)
from .ops import Act


class _ParallelHeadNet(nn.Module):
    """Shared passes; subclasses build the module tree in the reference's order and format the outputs."""

    gated = False
    gating_level = -1

    # stem, weight re-pack: the ResnestUNet implementations only touch modules both trees have
    _stem_fwd = ResnestUNet._stem_fwd
    _stem_bwd = ResnestUNet._stem_bwd
    _repack = ResnestUNet._repack

    def _build_trunk(self, num_classes: int, pretrain: bool, weight_path: Optional[str], gated: bool) -> None:
        resnest = resnest50(pretrained=pretrain, model_path=weight_path)
        self.num_classes = num_classes
        self.encoder_0_1_2 = nn.Sequential(resnest.conv1, resnest.bn1, resnest.relu)
        self.encoder_0_2_2 = resnest.maxpool
        self.upsampling_0 = Upsampling(64, 64)
        self.decoder_0 = ResNestDecoder(64, 32)
        if gated:
            self.aag_0 = AdversarialAttentionGate(32, num_classes)
        self.encoder_1 = resnest.layer1
        self.upsampling_1 = Upsampling(256, 64)
        self.decoder_1 = ResNestDecoder(128, 64)
        if gated:
            self.aag_1 = AdversarialAttentionGate(64, num_classes)
        self.encoder_2 = resnest.layer2
        self.upsampling_2 = Upsampling(512, 256)
        self.decoder_2 = ResNestDecoder(512, 256)
        if gated:
            self.aag_2 = AdversarialAttentionGate(256, num_classes)
        self.encoder_3 = resnest.layer3
        self.upsampling_3 = Upsampling(1024, 512)
        self.decoder_3 = ResNestDecoder(1024, 512)
        if gated:
            self.aag_3 = AdversarialAttentionGate(512, num_classes)
        self.encoder_4 = resnest.layer4
        self.upsampling_4 = Upsampling(2048, 1024)
        self.decoder_4 = ResNestDecoder(2048, 1024)
        if gated:
            self.aag_4 = AdversarialAttentionGate(1024, num_classes)
        self.upsampling_1_c = Upsampling(256, 64)
        self.decoder_1_c = ResNestDecoder(128, 64)
        if gated:
            self.aag_1_c = AdversarialAttentionGate(64, num_classes)
        self.upsampling_0_c = Upsampling(64, 64)
        self.decoder_0_c = ResNestDecoder(64, 32)
        if gated:
            self.aag_0_c = AdversarialAttentionGate(32, num_classes)
        self.fc = Conv2d(in_channels=32, out_channels=num_classes, kernel_size=1, stride=1)
        self.fc_c = Conv2d(in_channels=32, out_channels=num_classes, kernel_size=1, stride=1)
        self._maxpool_pd = ops.pool_desc("max", 3, 2, 1, False, True)
        self._grad_ready_hook = None

    def _hot_params(self) -> List[nn.Parameter]:
        return list(self.parameters())

    def _gate_on(self, level: int) -> bool:
        """compose.py:465-489: level 4 is gated for gating_level > 3, the others for gating_level >= level"""
        return self.gated and self.gating_level >= level

    # ---- explicit passes ---------------------------------------------------------------------------
    def _level(self, tape, name: str, inp: Act, on: bool, sink: list) -> Act:
        d, tape["dec" + name] = getattr(self, "decoder_" + name).fwd(inp)
        tape["aag" + name] = None
        if on:
            d, y, tape["aag" + name] = getattr(self, "aag_" + name).fwd(d)
            sink.append(y)
        return d

    def _fwd(self, x: Tensor):
        """x: [B,3,H,W] -> ((att..., att_c..., agg_map, agg_map_c) fp32 NCHW, tape)"""
        if not x.is_cuda:
            raise RuntimeError("octave_b200: input is on CPU; the B200 kernels have no CPU fallback")
        B, Cin, H, W = x.shape
        if Cin != 3:
            raise ValueError("the ResNeSt stem expects 3 input channels (resnest.py:327)")
        if H % 16 or W % 16:
            raise ValueError(f"input extent {H}x{W} must be a multiple of 16 (the reference's skip concatenations "
                             f"fail otherwise, compose.py:316-337)")
        dt, dev = compute_dtype(), x.device
        tape = {}
        if dt == torch.bfloat16:
            self._repack()
        xa = ops.nchw_to_nhwc(x, dt) if dt != torch.bfloat16 else x.detach()
        cat1 = Act.empty(B, H // 2, W // 2, 128, dt, dev)
        cat1c = Act.empty(B, H // 2, W // 2, 128, dt, dev)
        cat2 = Act.empty(B, H // 4, W // 4, 512, dt, dev)
        cat3 = Act.empty(B, H // 8, W // 8, 1024, dt, dev)
        cat4 = Act.empty(B, H // 16, W // 16, 2048, dt, dev)
        x_0_0, tape["stem"] = self._stem_fwd(xa, cat1.slice(0, 64))
        ops.copy_window(x_0_0, cat1c.slice(0, 64), False)          # the second concat of the same skip (compose.py:339)
        x_0_1, tape["maxpool"] = ops.pool_fwd(self._maxpool_pd, x_0_0)
        x_1, tape["enc1"] = layer_fwd(self.encoder_1, x_0_1, cat2.slice(0, 256))
        x_2, tape["enc2"] = layer_fwd(self.encoder_2, x_1, cat3.slice(0, 512))
        x_3, tape["enc3"] = layer_fwd(self.encoder_3, x_2, cat4.slice(0, 1024))
        h3, w3 = x_3.H, x_3.W
        if (h3 % 2) or (w3 % 2):
            x_3p = Act.empty(B, h3 + h3 % 2, w3 + w3 % 2, 1024, dt, dev)
            ops.copy_window(x_3, x_3p, False)
        else:
            x_3p = x_3
        tape["x3"] = (x_3, x_3p)
        x_4, tape["enc4"] = layer_fwd(self.encoder_4, x_3p)
        atts, atts_c = [], []
        _, tape["up4"] = self.upsampling_4.fwd(x_4, cat4.slice(1024, 1024))
        d = self._level(tape, "4", cat4, self._gate_on(4), atts)
        _, tape["up3"] = self.upsampling_3.fwd(d, cat3.slice(512, 512))
        d = self._level(tape, "3", cat3, self._gate_on(3), atts)
        _, tape["up2"] = self.upsampling_2.fwd(d, cat2.slice(256, 256))
        d = self._level(tape, "2", cat2, self._gate_on(2), atts)
        _, tape["up1"] = self.upsampling_1.fwd(d, cat1.slice(64, 64))
        d = self._level(tape, "1", cat1, self._gate_on(1), atts)
        up0, tape["up0"] = self.upsampling_0.fwd(d)
        d_0 = self._level(tape, "0", up0, self._gate_on(0), atts)
        # parallel branch (compose.py:338-343)
        _, tape["up1_c"] = self.upsampling_1_c.fwd(x_1, cat1c.slice(64, 64))
        dc = self._level(tape, "1_c", cat1c, self._gate_on(1), atts_c)
        up0c, tape["up0_c"] = self.upsampling_0_c.fwd(dc)
        d_0_c = self._level(tape, "0_c", up0c, self._gate_on(0), atts_c)
        heads = []
        for fc, dd, key in ((self.fc, d_0, "fc"), (self.fc_c, d_0_c, "fc_c")):
            w = fc.weight.detach().reshape(self.num_classes, -1).float().contiguous()
            b = fc.bias.detach().float().contiguous()
            agg, _ = ops.head_fwd(dd, w, b, 0)
            tape[key] = (dd, w, b)
            heads.append(agg)
        atts.reverse(); atts_c.reverse()
        tape["n_att"], tape["n_att_c"] = len(atts), len(atts_c)
        tape["dims"] = (B, H, W)
        return (*atts, *atts_c, heads[0], heads[1]), tape

    def _bwd(self, tape, gouts, grads: Grads) -> None:
        na, nc = tape["n_att"], tape["n_att_c"]
        g_att = list(gouts[:na]); g_att.reverse()                 # coarse level first, like the forward order
        g_att_c = list(gouts[na:na + nc]); g_att_c.reverse()
        g_agg, g_agg_c = gouts[na + nc], gouts[na + nc + 1]
        gy = dict(zip([l for l in (4, 3, 2, 1, 0) if self._gate_on(l)], g_att))
        gyc = dict(zip([l for l in (1, 0) if self._gate_on(l)], g_att_c))
        B, H, W = tape["dims"]
        dt = compute_dtype()

        def cont(t):
            return None if t is None else t.contiguous().float()

        def head_bwd(fc, key, g):
            dd, w, b = tape[key]
            if g is None:
                return None
            dx, dw, db = ops.head_bwd(dd, w, b, 0, cont(g), None)
            _acc(grads, fc.weight, dw); _acc(grads, fc.bias, db)
            return dx

        def level_bwd(name: str, dd: Optional[Act], gyh, like: Act) -> Optional[Act]:
            """gradient entering decoder_<name>'s output: through the gate when there is one"""
            ctx = tape["aag" + name]
            if ctx is not None and (dd is not None or gyh is not None):
                if dd is None:
                    dd = Act.zeros(like.B, like.H, like.W, like.C, dt, like.buf.device)
                dd = getattr(self, "aag_" + name).bwd(ctx, dd, cont(gyh), grads)
            if dd is None:
                return None
            return getattr(self, "decoder_" + name).bwd(tape["dec" + name], dd, grads)

        # ---- parallel branch first: its gradients wait at x_1 / x_0_0 for the main path
        d0c_out = tape["fc_c"][0]
        ddc = head_bwd(self.fc_c, "fc_c", g_agg_c)
        dup0c = level_bwd("0_c", ddc, gyc.get(0), d0c_out)
        dx1_c = dcat1c = None
        ddc = self.upsampling_0_c.bwd(tape["up0_c"], dup0c, grads) if dup0c is not None else None
        d1c_out = tape["up0_c"][0]
        dcat1c = level_bwd("1_c", ddc, gyc.get(1), d1c_out)
        if dcat1c is not None:
            dx1_c = self.upsampling_1_c.bwd(tape["up1_c"], dcat1c.slice(64, 64), grads)
        # ---- main path
        dd = head_bwd(self.fc, "fc", g_agg)
        dup0 = level_bwd("0", dd, gy.get(0), tape["fc"][0])
        dd = self.upsampling_0.bwd(tape["up0"], dup0, grads) if dup0 is not None else None
        dcat1 = level_bwd("1", dd, gy.get(1), tape["up0"][0])
        dd = self.upsampling_1.bwd(tape["up1"], dcat1.slice(64, 64), grads) if dcat1 is not None else None
        dcat2 = level_bwd("2", dd, gy.get(2), tape["up1"][0])
        dd = self.upsampling_2.bwd(tape["up2"], dcat2.slice(256, 256), grads) if dcat2 is not None else None
        dcat3 = level_bwd("3", dd, gy.get(3), tape["up2"][0])
        dd = self.upsampling_3.bwd(tape["up3"], dcat3.slice(512, 512), grads) if dcat3 is not None else None
        dcat4 = level_bwd("4", dd, gy.get(4), tape["up3"][0])
        # ---- encoder: every stage sums the gradient from above with its skip gradients
        x_3, x_3p = tape["x3"]
        dx3 = None
        if dcat4 is not None:
            dx4 = self.upsampling_4.bwd(tape["up4"], dcat4.slice(1024, 1024), grads)
            dx3p = layer_bwd(self.encoder_4, tape["enc4"], dx4, grads)
            dx3 = dcat4.slice(0, 1024)
            ops.copy_window(dx3p, dx3, True)
        dx2 = _sum_into(layer_bwd(self.encoder_3, tape["enc3"], dx3, grads) if dx3 is not None else None,
                        dcat3.slice(0, 512) if dcat3 is not None else None)
        dx1 = layer_bwd(self.encoder_2, tape["enc2"], dx2, grads) if dx2 is not None else None
        dx1 = _sum_into(dx1, dcat2.slice(0, 256) if dcat2 is not None else None)
        dx1 = _sum_into(dx1, dx1_c)
        dx00 = None
        if dx1 is not None:
            dx01 = layer_bwd(self.encoder_1, tape["enc1"], dx1, grads)
            dx00 = ops.pool_bwd(self._maxpool_pd, dx01, tape["maxpool"], H // 2, W // 2)
        dx00 = _sum_into(dx00, dcat1.slice(0, 64) if dcat1 is not None else None)
        dx00 = _sum_into(dx00, dcat1c.slice(0, 64) if dcat1c is not None else None)
        if dx00 is not None:
            self._stem_bwd(tape["stem"], dx00, grads)
        _n.join_side_stream(grads)
        hook = self._grad_ready_hook
        if hook is not None and grads:
            hook(list(grads.keys()), list(grads.values()))

    def _run(self, x: Tensor):
        if not self.training and not torch.is_grad_enabled() and config.fold_bn_inference:
            _n._fold_active = True
            try:
                outs, _ = self._fwd(x)
            finally:
                _n._fold_active = False
        else:
            outs = _n._SegmentorFn.apply(self, x, *self._hot_params())
        return outs


def _sum_into(a: Optional[Act], b: Optional[Act]) -> Optional[Act]:
    """a += b on Acts that may be absent; a slice view of a concat-gradient buffer is copied before it is written"""
    if b is None:
        return a
    if a is None:
        a = Act.empty(b.B, b.H, b.W, b.C, b.dtype, b.buf.device)
        ops.copy_window(b, a, False)
        return a
    ops.add_inplace(a, b)
    return a


class ResnestUnetParallelHead(_ParallelHeadNet):
    """reference: segmentor/compose.py:233-361.  forward -> agg [2, B, num_classes, H, W] (main head, parallel head)."""

    def __init__(self, num_classes: int, pretrain: bool, weight_path: str = None):
        super().__init__()
        self._build_trunk(num_classes, pretrain, weight_path, gated=False)

    def forward(self, x) -> Tensor:
        outs = self._run(x)
        return torch.stack([outs[-2], outs[-1]])                                   # compose.py:350

    def predict(self, x: Tensor, method='softmax'):
        """reference: compose.py:352-361"""
        agg_map = self.forward(x)
        if method == 'softmax':
            predicate = nn.Softmax(dim=2)(agg_map)
        elif method == 'sigmoid':
            predicate = nn.Sigmoid()(agg_map)
        elif method == 'one-hot':
            predicate = torch.nn.functional.one_hot(torch.argmax(agg_map, dim=2)).permute(0, 1, 4, 2, 3)
        elif method == 'original':
            predicate = agg_map
        return predicate


class ResnestUnetParallelHeadAttentionGate(_ParallelHeadNet):
    """reference: segmentor/compose.py:364-527 (the constructor keyword really is `gating_leveL`).
    forward -> ((attentions), (attentions_c)), agg [2, B, num_classes, H, W]; both tuples full-resolution first."""

    gated = True

    def __init__(self, num_classes: int, pretrain: bool, weight_path: str = None, gating_leveL: int = 3):
        super().__init__()
        self.gating_level = gating_leveL
        self._build_trunk(num_classes, pretrain, weight_path, gated=True)

    def _gate_on(self, level: int) -> bool:
        return self.gating_level > 3 if level == 4 else self.gating_level >= level

    def forward(self, x):
        outs = self._run(x)
        na = sum(1 for l in (4, 3, 2, 1, 0) if self._gate_on(l))
        nc = sum(1 for l in (1, 0) if self._gate_on(l))
        return (tuple(outs[:na]), tuple(outs[na:na + nc])), torch.stack([outs[-2], outs[-1]])   # compose.py:515

    def predict(self, x: Tensor, method='softmax'):
        """reference: compose.py:517-527"""
        attentions, agg_map = self.forward(x)
        if method == 'softmax':
            predicate = nn.Softmax(dim=2)(agg_map)
        elif method == 'sigmoid':
            predicate = nn.Sigmoid()(agg_map)
        elif method == 'one-hot':
            predicate = torch.nn.functional.one_hot(torch.argmax(agg_map, dim=2)).permute(0, 1, 4, 2, 3)
        elif method == 'original':
            predicate = agg_map
        return attentions, predicate
