"""Multi-scale LSGAN mask critic on the sm_100a kernels (reference: /root/reference/architectures/discriminator/blocks.py:12-185).

Module tree, parameter names (incl. the legacy spectral-norm `weight_orig/_u/_v`) and init order mirror the reference
so that state_dicts are interchangeable and a seeded construction gives identical weights.

Noise handling: the reference draws `torch.normal(size=(H,W))` (InstanceNoise, :150) and `FloatTensor(1).uniform_()`
(LabelNoise via utils.rand_uniform, utils.py:20-22) on the CPU generator per call, moves them to the GPU and — for the
label flip — reads the comparison back (a device sync).  Here the same CPU draws are made in the same order, the noise
plane is uploaded asynchronously and the flip decision is taken on the host: same RNG stream, no device sync.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn, Tensor
from torch.nn.init import kaiming_normal_, xavier_uniform_
from torch.nn.utils import spectral_norm

from . import ops
from .network import Grads, _acc, compute_dtype
from .ops import Act, ConvSpec


def rand_uniform(x: Optional[Tensor] = None):
    """reference utils.py:20-22: one CPU uniform draw, returned `type_as(x)` (x=None raises there too).  LabelNoise below
    calls `_rand_uniform_host()` instead so that the comparison with `prob` needs no device read-back."""
    return _rand_uniform_host().type_as(x)


def _rand_uniform_host() -> Tensor:
    return torch.FloatTensor(1).uniform_(0, 1)


class InstanceNoise(nn.Module):
    def __init__(self, input_shape, mean: float, std: float, clipping: bool, is_training: bool):
        super().__init__()
        self.mean, self.std, self.clipping = mean, std, clipping
        self.size = (input_shape[2], input_shape[3])
        self.is_training = is_training

    def draw(self) -> Tensor:
        return torch.normal(mean=self.mean, std=self.std, size=self.size)   # blocks.py:150 (CPU generator)


class LabelNoise(nn.Module):
    def __init__(self, prob: float = 0.1, mode: str = 'sign'):
        super().__init__()
        self.prob, self.mode = prob, mode
        if mode != 'sign':
            raise NotImplementedError("octave_b200: LabelNoise mode 'label' is not constructed by DiscriminatorBlock (blocks.py:77)")

    def draw_flip(self) -> bool:
        return bool(_rand_uniform_host() < self.prob)                         # blocks.py:165-167 (same CPU draw)


class HostRandomFeed:
    """Static pinned + device buffers for the per-call CPU draws of InstanceNoise / LabelNoise.  With a feed installed
    (`DiscriminatorBlock._rand_feed`) a forward call takes its noise plane and label-flip sign from slot k of these
    buffers instead of drawing: `draw()` makes the very same CPU draws in the same order for all the calls of one step,
    `upload()` is the stream-ordered host-to-device copy issued before each replay.  A captured CUDA graph of the training step therefore sees fresh
    host randomness at every replay while the CPU generator advances exactly as in eager mode."""

    RING = 2      # pinned staging slots: the host fills slot i+1 while the upload of slot i may still be pending

    def __init__(self, disc: "DiscriminatorBlock", n_calls: int, device):
        self.disc, self.n_calls, self.slot = disc, n_calls, 0
        inst = disc._inst
        self.has_noise = inst is not None and inst.is_training
        if self.has_noise:
            self.noise_pin = [[torch.empty(inst.size, dtype=torch.float32).pin_memory() for _ in range(n_calls)] for _ in range(self.RING)]
            self.noise_dev = [torch.empty(inst.size, dtype=torch.float32, device=device) for _ in range(n_calls)]
        self.sign_pin = [torch.ones(n_calls, dtype=torch.float32).pin_memory() for _ in range(self.RING)]
        self.sign_dev = torch.ones(n_calls, dtype=torch.float32, device=device)
        self._ring = 0
        self._uploaded = [None] * self.RING          # event recorded after the upload that reads ring slot r

    def draw(self) -> None:
        """The CPU draws of one step (same generator, same order as the eager modules) into the next pinned ring slot.
        Blocks only if the upload issued RING steps ago from this slot has not run yet."""
        d = self.disc
        self._ring = (self._ring + 1) % self.RING
        ev = self._uploaded[self._ring]
        if ev is not None:
            ev.synchronize()
        for k in range(self.n_calls):
            if d._inst is not None:
                noise = d._inst.draw()
                if self.has_noise:
                    self.noise_pin[self._ring][k].copy_(noise)
            if d._label is not None:
                self.sign_pin[self._ring][k] = -1.0 if d._label.draw_flip() else 1.0

    def upload(self) -> None:
        """Asynchronous host-to-device copy of the slot filled by the last draw() (stream-ordered: NOT part of a captured
        graph, so that the pinned source can rotate)."""
        r = self._ring
        if self.has_noise:
            for k in range(self.n_calls):
                self.noise_dev[k].copy_(self.noise_pin[r][k], non_blocking=True)
        self.sign_dev.copy_(self.sign_pin[r], non_blocking=True)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("octave_b200: HostRandomFeed.upload() must not be captured")
        ev = self._uploaded[r] or torch.cuda.Event()
        ev.record()
        self._uploaded[r] = ev
        self.slot = 0

    def rewind(self) -> None:
        """start of a step whose buffers were uploaded outside (graph replay): calls take slots 0.. again"""
        self.slot = 0

    def next(self):
        k = self.slot
        if k >= self.n_calls:
            raise RuntimeError("octave_b200: more discriminator calls in one step than the random feed was sized for")
        self.slot += 1
        return (self.noise_dev[k] if self.has_noise else None), self.sign_dev[k:k + 1]


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: "DiscriminatorBlock", n_maps: int, *args):
        maps, params = args[:n_maps], args[n_maps:]
        logits, tape = net._fwd(maps)
        ctx.net, ctx.tape, ctx.params, ctx.n_maps = net, tape, params, n_maps
        ctx.set_materialize_grads(False)
        return logits

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        need_params = any(ctx.needs_input_grad[2 + ctx.n_maps:])
        need_maps = any(ctx.needs_input_grad[2:2 + ctx.n_maps])
        grads: Grads = {}
        gmaps = net._bwd(ctx.tape, g, grads, need_params, need_maps)
        ctx.tape = None
        return (None, None, *gmaps, *[grads.get(p) for p in ctx.params])


class DiscriminatorBlock(nn.Module):
    def __init__(self, input_shape, is_training: bool, depth: int = 3, num_filters: int = 64,
                 instance_noise: bool = True, label_noise: bool = True):
        super().__init__()
        self.num_filters, self.is_training, self.depth = num_filters, is_training, depth
        in_channels = input_shape[1]
        self.in_channels = in_channels
        modules = []
        if instance_noise:
            modules.append(InstanceNoise(input_shape=input_shape, is_training=is_training, mean=.0, std=.2, clipping=True))
        conv_0 = nn.Conv2d(in_channels, num_filters, kernel_size=4, stride=2, padding=1)
        kaiming_normal_(conv_0.weight, nonlinearity='leaky_relu')
        modules.append(conv_0)
        modules.append(nn.LeakyReLU(negative_slope=0.2))
        self.stack_0 = nn.Sequential(*modules)
        squeeze_stack, spectral_stack = dict(), dict()
        for i in range(self.depth):
            squeeze = nn.Sequential(nn.Conv2d(num_filters * (2 ** i), 13, kernel_size=1, stride=1), nn.Sigmoid())
            conv = nn.Conv2d(13 + in_channels, num_filters * 2 * (2 ** i), kernel_size=4, stride=2, padding=1)
            conv = spectral_norm(conv, n_power_iterations=1)
            squeeze_stack[f'squeeze_{i}'] = squeeze
            spectral_stack[f'spectral_{i}'] = nn.Sequential(conv, nn.Tanh())
        self.squeeze_dict = nn.ModuleDict(squeeze_stack)
        self.spectral_dict = nn.ModuleDict(spectral_stack)
        h, w = [int(i) // (2 ** (self.depth + 1)) for i in input_shape[2:]]
        fc = nn.Conv2d(num_filters * (2 ** self.depth), out_channels=1, kernel_size=(h, w), stride=1)
        xavier_uniform_(fc.weight)
        modules = [fc, nn.Flatten()]
        if label_noise:
            modules.append(LabelNoise(0.1, 'sign'))
        self.out = nn.Sequential(*modules)
        self._has_inst, self._has_label = bool(instance_noise), bool(label_noise)

    @property
    def _inst(self):
        return self.stack_0[0] if self._has_inst else None

    @property
    def _conv0(self):
        return self.stack_0[1 if self._has_inst else 0]

    @property
    def _label(self):
        return self.out[2] if self._has_label else None

    # ---- spectral norm (torch.nn.utils.spectral_norm legacy semantics, n_power_iterations=1, eps=1e-12) ----
    def _spectral_weight(self, conv: nn.Module, need_w: bool = True):
        """-> (W_orig / sigma or None, sigma, u, v) after one power iteration in training mode."""
        w = conv.weight_orig.detach()
        wm = w.reshape(w.shape[0], -1)
        if wm.dtype != torch.float32 or not wm.is_contiguous():
            wm = wm.float().contiguous()
        u, v = conv.weight_u, conv.weight_v
        if wm.is_cuda and wm.shape[0] <= 1024 and u.dtype == torch.float32 and v.dtype == torch.float32:
            # one launch: power iteration (u, v in place) + sigma + 1/sigma, fixed summation order
            sg = ops.spectral_sigma(wm, u, v, self.training)
            sigma = sg[0]
            self._last_inv_sigma = sg[1:2]
            return (w.float() / sigma if need_w else None), sigma, u.clone(), v.clone()
        if self.training:
            v_new = ops.glinear_bwd_data_only(u.reshape(1, -1), wm)          # W^T u
            v_new = v_new / v_new.norm().clamp_min(1e-12)
            wv = ops.glinear_fwd(v_new, wm, None, 1, 1.0)                     # W v
            u_new = wv / wv.norm().clamp_min(1e-12)
            u.copy_(u_new.reshape(-1)); v.copy_(v_new.reshape(-1))
            wv = wv.reshape(-1)                                               # = W v for the updated v: reused for sigma
        else:
            wv = ops.glinear_fwd(v.reshape(1, -1).contiguous(), wm, None, 1, 1.0).reshape(-1)
        sigma = (u * wv).sum()
        self._last_inv_sigma = (1.0 / sigma).reshape(1).float()
        return (w.float() / sigma if need_w else None), sigma, u.clone(), v.clone()

    @staticmethod
    def _spectral_wgrad(dwsn: Tensor, sn: nn.Module, sigma: Tensor, u: Tensor, v: Tensor) -> Tensor:
        """W = W_orig / sigma, sigma = u^T W_orig v (u, v constants): dW_orig = dW/sigma - <dW, W_orig>/sigma^2 * u v^T."""
        wo = sn.weight_orig.detach()
        if (dwsn.is_cuda and dwsn.dtype == torch.float32 and wo.dtype == torch.float32 and dwsn.is_contiguous() and wo.is_contiguous()
                and dwsn.shape == wo.shape and u.dtype == torch.float32 and v.dtype == torch.float32 and sigma.dtype == torch.float32
                and u.is_contiguous() and v.is_contiguous()):
            return ops.spectral_wgrad(dwsn, wo, u, v, sigma)           # two launches instead of nine torch ops
        wo = wo.float()
        coef = (dwsn * wo).sum() / (sigma * sigma)
        uv = (u.reshape(-1, 1) * v.reshape(1, -1)).reshape(wo.shape)
        return dwsn / sigma - coef * uv

    def _spectral_all(self):
        """Every spectral-norm layer of this critic call at once -> per level (sigma, inv_sigma, u, v): ONE launch instead of
        one per level in front of its conv (the power iterations depend on the weights only), u / v snapshots included."""
        sns = [self.spectral_dict[f'spectral_{i}'][0] for i in range(self.depth)]
        ws = [sn.weight_orig.detach().reshape(sn.weight_orig.shape[0], -1) for sn in sns]
        ok = 0 < len(sns) <= ops.SN_MAX_JOBS and all(
            w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.shape[0] <= 1024 and w.shape[1] <= 8192
            and sn.weight_u.dtype == torch.float32 and sn.weight_v.dtype == torch.float32 for w, sn in zip(ws, sns))
        if not ok:
            return None
        outs = ops.spectral_sigma_multi(ws, [sn.weight_u for sn in sns], [sn.weight_v for sn in sns], self.training)
        res = []
        for w, o in zip(ws, outs):
            r, c = w.shape
            res.append((o[0], o[1:2], o[2:2 + r], o[2 + r:2 + r + c]))
        return res

    def _padded_squeeze(self, i: int, rows: int):
        """13-row squeeze conv weight / bias zero-padded to `rows` output channels for the tensor-core kernels; rebuilt
        only when the parameter changes (once per optimiser step, not once per critic call)."""
        sq = self.squeeze_dict[f'squeeze_{i}'][0]
        cache = self.__dict__.setdefault("_sq_cache", {})
        key = (sq.weight._version, sq.weight.data_ptr(), sq.bias._version, sq.bias.data_ptr())
        hit = cache.get((i, rows))
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        dev = sq.weight.device
        wpad = torch.zeros((rows, sq.in_channels, 1, 1), dtype=torch.float32, device=dev)
        wpad[:13] = sq.weight.detach()
        bpad = torch.zeros(rows, dtype=torch.float32, device=dev)
        bpad[:13] = sq.bias.detach()
        spec = ConvSpec(wpad, None, sq.in_channels, rows, 1, 1, 0, 1)
        cache[(i, rows)] = (key, spec, bpad)
        return spec, bpad

    # ---- explicit passes -------------------------------------------------------------------------------
    def _use_tc(self) -> bool:
        return compute_dtype() == torch.bfloat16 and self.in_channels <= 3 and self.num_filters % 32 == 0

    def _fwd(self, maps: Sequence[Tensor]):
        if self._use_tc():
            return self._fwd_tc(maps)
        return self._fwd_plain(maps)

    def _bwd(self, tape, g, grads, need_params, need_maps):
        if tape.get("tc"):
            return self._bwd_tc(tape, g, grads, need_params, need_maps)
        return self._bwd_plain(tape, g, grads, need_params, need_maps)

    # ---- tensor-core path: every 4x4 s2 conv becomes a 3x3 conv over a space-to-depth input -----------------
    def _noise_and_flip(self, dev):
        """-> (noise plane on the device or None, clip flag, flip): flip is a host bool, or — with a HostRandomFeed
        installed — a one-element device tensor holding the sign (+1 / -1) the logits are multiplied with."""
        feed = getattr(self, "_rand_feed", None)
        clip = self._inst.clipping if self._inst is not None else False
        if feed is not None:
            noise_dev, sign = feed.next()
            return noise_dev, clip, (sign if self._label is not None else False)
        noise_dev = None
        if self._inst is not None:
            noise = self._inst.draw()
            if self._inst.is_training:
                noise_dev = noise.pin_memory().to(dev, non_blocking=True)
        flip = self._label.draw_flip() if self._label is not None else False
        return noise_dev, clip, flip

    def _fwd_tc(self, maps: Sequence[Tensor]):
        if len(maps) < self.depth + 1:
            raise Exception(f'Exception raised in depth = {len(maps) - 1}')
        y0 = maps[0]
        if not y0.is_cuda:
            raise RuntimeError("octave_b200: discriminator input is on CPU; the B200 kernels have no CPU fallback")
        dt, dev = torch.bfloat16, y0.device
        B, Cin, H, W = y0.shape
        noise_dev, clip, flip = self._noise_and_flip(dev)
        tape = {"tc": True}
        y0c = y0.detach().contiguous().float()
        X0 = Act.empty(B, (H + 1) // 2, (W + 1) // 2, 32, dt, dev)                         # quadrant stride 8; nchw_to_s2d writes every channel
        ops.nchw_to_s2d(y0c, X0, 8, 0, noise_dev, clip)                                   # blocks.py:149-154
        c0 = self._conv0
        nf = c0.out_channels
        ho, wo = (H + 2 - 4) // 2 + 1, (W + 2 - 4) // 2 + 1
        s = ops.conv4x4s2_tc_fwd(X0, ops.pack_weight_s2d(c0.weight.detach(), None, 0, 8), c0.bias.detach().float(), nf, ho, wo,
                                 ops.ACT_LEAKY)                                           # :46-50
        tape["in"] = (y0c, noise_dev, clip, X0, s, (H, W))
        levels = []
        sn_all = self._spectral_all()
        for i in range(self.depth):
            sq = self.squeeze_dict[f'squeeze_{i}'][0]
            sn = self.spectral_dict[f'spectral_{i}'][0]
            yi = maps[i + 1]
            if yi.shape[2] != s.H or yi.shape[3] != s.W:
                raise Exception(f'Exception raised in depth = {i}')
            h, w = s.H, s.W
            # quadrant stride 16: 13 + Cin (+pad).  Even sizes: the squeeze conv and nchw_to_s2d write every channel of every
            # pixel; odd sizes leave the quadrant pixels outside the map to the memset
            catS = (Act.zeros if (h | w) & 1 else Act.empty)(B, (h + 1) // 2, (w + 1) // 2, 64, dt, dev)
            spec_sq, b16 = self._padded_squeeze(i, 16)
            ops.conv1x1_tc_s2d_store(s, spec_sq.pack(0), b16, 16, catS, 16, ops.ACT_SIGMOID)   # :121
            ops.nchw_to_s2d(yi.detach(), catS, 16, 13)                                    # :122 (overwrites pad channels 13,14)
            if sn_all is not None:
                sigma, inv_sigma, u, v = sn_all[i]
            else:
                _, sigma, u, v = self._spectral_weight(sn, need_w=False)
                inv_sigma = self._last_inv_sigma
            wo_ = sn.weight_orig.detach()
            ho, wo = (h + 2 - 4) // 2 + 1, (w + 2 - 4) // 2 + 1
            s_next = ops.conv4x4s2_tc_fwd(catS, ops.pack_weight_s2d(wo_, inv_sigma, 0, 16), sn.bias.detach().float(),
                                          sn.out_channels, ho, wo, ops.ACT_TANH)          # :123
            levels.append((s, catS, sigma, inv_sigma, u, v, s_next, (h, w)))
            s = s_next
        fc = self.out[0]
        kh, kw = fc.kernel_size
        if (s.H, s.W) != (kh, kw):
            raise RuntimeError(f"octave_b200: final feature map {s.H}x{s.W} != output kernel {kh}x{kw}")
        w_hwc = fc.weight.detach().float().permute(0, 2, 3, 1).reshape(-1).contiguous()
        logits = ops.rowdot_fwd(s, w_hwc, fc.bias.detach().float())
        if isinstance(flip, Tensor):
            logits = logits * flip                                                        # :167-168, sign as data
        elif flip:
            logits = -1 * logits
        tape["levels"], tape["out"], tape["flip"] = levels, (s, w_hwc), flip
        return logits, tape

    def _bwd_tc(self, tape, g: Tensor, grads: Grads, need_params: bool, need_maps: bool):
        g = g.contiguous().float()
        if isinstance(tape["flip"], Tensor):
            g = g * tape["flip"]
        elif tape["flip"]:
            g = -1 * g
        s_last, w_hwc = tape["out"]
        fc = self.out[0]
        ds, dw, db = ops.rowdot_bwd(s_last, w_hwc, g.reshape(-1), need_params)
        if need_params:
            kh, kw = fc.kernel_size
            _acc(grads, fc.weight, dw.reshape(1, kh, kw, -1).permute(0, 3, 1, 2).contiguous())
            _acc(grads, fc.bias, db)
        gmaps: List[Optional[Tensor]] = [None] * (self.depth + 1)
        Cin = self.in_channels
        for i in reversed(range(self.depth)):
            s_in, catS, sigma, inv_sigma, u, v, s_out, (h, w) = tape["levels"][i]
            sq = self.squeeze_dict[f'squeeze_{i}'][0]
            sn = self.spectral_dict[f'spectral_{i}'][0]
            dz = ops.act_bwd(s_out, ds, ops.ACT_TANH, out=ds)
            if need_params:
                dwsn = ops.conv4x4s2_tc_wgrad(catS, dz, 13 + Cin, 16)
                _acc(grads, sn.weight_orig, self._spectral_wgrad(dwsn, sn, sigma, u, v))
                _acc(grads, sn.bias, ops.chan_stats(dz)[:dz.C].float())
            dcatS = ops.conv4x4s2_tc_dgrad(dz, ops.pack_weight_s2d(sn.weight_orig.detach(), inv_sigma, 1, 16), catS.H, catS.W, 64)
            if need_maps:
                gmaps[i + 1] = ops.s2d_to_nchw(dcatS, 16, 13, Cin, h, w)
            ops.act_bwd(catS, dcatS, ops.ACT_SIGMOID, out=dcatS)
            # squeeze-conv backward on the tensor cores: the 13 gradient channels are padded to 32 (channels 13..15 hold
            # finite don't-care values of the mask / pad quadrant channels, 16..31 zeros) and meet zero weight rows
            dsq32 = Act.zeros(catS.B, h, w, 32, catS.dtype, catS.device)
            ops._chk("octave_depth_to_space", ops.lib.octave_depth_to_space(ops._ref(dcatS), ops._ref(dsq32.slice(0, 16)), ops.stream_ptr()))
            spec_sq, _ = self._padded_squeeze(i, 32)
            if need_params:
                dwq, _ = ops.conv_wgrad(s_in, dsq32, spec_sq)
                _acc(grads, sq.weight, dwq[:13].contiguous())
                _acc(grads, sq.bias, ops.chan_stats(dsq32)[:13].float())
            ds = ops.conv_dgrad(dsq32, spec_sq, s_in.H, s_in.W)
        y0c, noise_dev, clip, X0, s0, (H, W) = tape["in"]
        dz0 = ops.act_bwd(s0, ds, ops.ACT_LEAKY, out=ds)
        c0 = self._conv0
        if need_params:
            _acc(grads, c0.weight, ops.conv4x4s2_tc_wgrad(X0, dz0, Cin, 8))
            _acc(grads, c0.bias, ops.chan_stats(dz0)[:dz0.C].float())
        if need_maps:
            dX0 = ops.conv4x4s2_tc_dgrad(dz0, ops.pack_weight_s2d(c0.weight.detach(), None, 1, 8), X0.H, X0.W, 32)
            gmaps[0] = ops.s2d_to_nchw(dX0, 8, 0, Cin, H, W, y0c, noise_dev, clip)
        return gmaps

    # ---- CUDA-core path (fp32 mode): plain NHWC, direct 4x4 s2 convs ----------------------------------------
    def _fwd_plain(self, maps: Sequence[Tensor]):
        if len(maps) < self.depth + 1:
            raise Exception(f'Exception raised in depth = {len(maps) - 1}')
        y0 = maps[0]
        if not y0.is_cuda:
            raise RuntimeError("octave_b200: discriminator input is on CPU; the B200 kernels have no CPU fallback")
        dt, dev = compute_dtype(), y0.device
        B, Cin, H, W = y0.shape
        tape = {}
        noise_dev, clip = None, False
        if self._inst is not None:
            noise = self._inst.draw()
            if self._inst.is_training:
                noise_dev = noise.pin_memory().to(dev, non_blocking=True)
            clip = self._inst.clipping
        flip = self._label.draw_flip() if self._label is not None else False
        x0 = Act.zeros(B, H, W, 8, dt, dev)
        x0v = x0.slice(0, Cin)
        y0c = y0.detach().contiguous().float()
        ops.nchw_into(y0c, x0v, noise_dev, clip)                                          # blocks.py:149-154
        c0 = self._conv0
        spec0 = ConvSpec(c0.weight, c0.bias, Cin, c0.out_channels, 4, 2, 1, 1)
        s = ops.conv_fwd(x0v, spec0, act=ops.ACT_LEAKY)                                   # :46-50
        tape["in"] = (y0c, noise_dev, clip, x0v, spec0, s)
        levels = []
        for i in range(self.depth):
            sq = self.squeeze_dict[f'squeeze_{i}'][0]
            sn = self.spectral_dict[f'spectral_{i}'][0]
            yi = maps[i + 1]
            if yi.shape[2] != s.H or yi.shape[3] != s.W:
                raise Exception(f'Exception raised in depth = {i}')
            cat = Act.zeros(B, s.H, s.W, 16, dt, dev)
            spec_sq = ConvSpec(sq.weight, sq.bias, sq.in_channels, 13, 1, 1, 0, 1)
            ops.conv_fwd(s, spec_sq, out=cat.slice(0, 13), act=ops.ACT_SIGMOID)           # :121
            ops.nchw_into(yi.detach(), cat.slice(13, Cin))                                # :122
            w_sn, sigma, u, v = self._spectral_weight(sn)
            spec_sn = ConvSpec(w_sn.contiguous(), sn.bias, 13 + Cin, sn.out_channels, 4, 2, 1, 1)
            catv = cat.slice(0, 13 + Cin)
            s_next = ops.conv_fwd(catv, spec_sn, act=ops.ACT_TANH)                        # :123
            levels.append((s, spec_sq, cat, catv, spec_sn, sigma, u, v, s_next))
            s = s_next
        fc = self.out[0]
        kh, kw = fc.kernel_size
        if (s.H, s.W) != (kh, kw):
            raise RuntimeError(f"octave_b200: final feature map {s.H}x{s.W} != output kernel {kh}x{kw}")
        w_hwc = fc.weight.detach().float().permute(0, 2, 3, 1).reshape(-1).contiguous()
        logits = ops.rowdot_fwd(s, w_hwc, fc.bias.detach().float())                       # :128, [B,1]
        if flip:
            logits = -1 * logits                                                          # :167-168
        tape["levels"], tape["out"], tape["flip"] = levels, (s, w_hwc), flip
        return logits, tape

    def _bwd_plain(self, tape, g: Tensor, grads: Grads, need_params: bool, need_maps: bool):
        g = g.contiguous().float()
        if tape["flip"]:
            g = -1 * g
        s_last, w_hwc = tape["out"]
        fc = self.out[0]
        ds, dw, db = ops.rowdot_bwd(s_last, w_hwc, g.reshape(-1), need_params)
        if need_params:
            kh, kw = fc.kernel_size
            _acc(grads, fc.weight, dw.reshape(1, kh, kw, -1).permute(0, 3, 1, 2).contiguous())
            _acc(grads, fc.bias, db)
        gmaps: List[Optional[Tensor]] = [None] * (self.depth + 1)
        Cin = self.in_channels
        for i in reversed(range(self.depth)):
            s_in, spec_sq, cat, catv, spec_sn, sigma, u, v, s_out = tape["levels"][i]
            sq = self.squeeze_dict[f'squeeze_{i}'][0]
            sn = self.spectral_dict[f'spectral_{i}'][0]
            dz = ops.act_bwd(s_out, ds, ops.ACT_TANH, out=ds)
            if need_params:
                dwsn, dbsn = ops.conv_wgrad(catv, dz, spec_sn)
                _acc(grads, sn.weight_orig, self._spectral_wgrad(dwsn, sn, sigma, u, v))
                _acc(grads, sn.bias, dbsn)
            dcat = Act.zeros(cat.B, cat.H, cat.W, 16, cat.dtype, cat.device)
            dcatv = dcat.slice(0, 13 + Cin)
            ops.conv_dgrad(dz, spec_sn, cat.H, cat.W, out=dcatv)
            if need_maps:
                gmaps[i + 1] = ops.nhwc_to_nchw(dcat.slice(13, Cin))
            dsq = ops.act_bwd(cat.slice(0, 13), dcat.slice(0, 13), ops.ACT_SIGMOID, out=dcat.slice(0, 13))
            if need_params:
                dwq, dbq = ops.conv_wgrad(s_in, dsq, spec_sq)
                _acc(grads, sq.weight, dwq); _acc(grads, sq.bias, dbq)
            ds = ops.conv_dgrad(dsq, spec_sq, s_in.H, s_in.W)
        y0c, noise_dev, clip, x0v, spec0, s0 = tape["in"]
        dz0 = ops.act_bwd(s0, ds, ops.ACT_LEAKY, out=ds)
        if need_params:
            dw0, db0 = ops.conv_wgrad(x0v, dz0, spec0)
            _acc(grads, self._conv0.weight, dw0); _acc(grads, self._conv0.bias, db0)
        if need_maps:
            dx0 = Act.zeros(x0v.B, x0v.H, x0v.W, 8, x0v.dtype, x0v.device).slice(0, Cin)
            ops.conv_dgrad(dz0, spec0, x0v.H, x0v.W, out=dx0)
            gmaps[0] = ops.nhwc_to_nchw_clipmask(dx0, y0c, noise_dev, clip)
        return gmaps

    def forward(self, y: Sequence[Tensor]):
        """y: list of multi-scale maps [B,C,H/2^k,W/2^k] -> logits [B,1] (reference blocks.py:114-130)."""
        maps = list(y[:self.depth + 1])
        if len(y) < self.depth + 1:
            raise Exception(f'Exception raised in depth = {len(y) - 1}')
        params = list(self.parameters())
        return _DiscFn.apply(self, len(maps), *maps, *params)

    def predict(self, y: List[Tensor]):
        return self.forward(y)
