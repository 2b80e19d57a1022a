"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32/fp64, functional, no einops/kornia) of the
reference's arithmetic on the hot path.  It is the checker for the CUDA kernels; the product never
imports it.  Every function cites the reference lines it restates (paths relative to /root/reference).

Pinned against the real reference by tests/test_oracle_vs_reference.py (runs where /root/reference is
mounted) and by the golden vectors in tests/golden/ generated with oracle/make_golden.py from the
imported reference itself.  The reference ships no tests or golden vectors of its own.

The network functions take a `state_dict`-style mapping with the reference's exact keys, so weights
initialised by either implementation can be fed to both.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

EPS = 1e-12


# =================================================================================================
# Losses
# =================================================================================================
def weighted_partial_ce(y_hat: Tensor, ys: Tensor, num_classes: int, reduction: str = "mean",
                        full: bool = False) -> Tensor:
    """architectures/segmentor/losses.py:26-61, manual=True branch (:51-55).  ignore_bg is the caller's
    in-place `ys[:,0]=0` (:29-30) and is applied by the caller before this function."""
    assert y_hat.shape[1] == ys.shape[1], 'Number of class mismatch.'
    if not full:
        y_hat = y_hat * ys                                    # :31-32
    ni = ys.sum(dim=(0, 2, 3))                                # :34
    n_tot = ni.sum()                                          # :36
    weights = torch.stack([n_tot / (ni[c] + 1e-12) for c in range(num_classes)])  # :37
    yh = y_hat.permute(0, 2, 3, 1).reshape(-1, y_hat.shape[1])  # :40
    yt = ys.permute(0, 2, 3, 1).reshape(-1, ys.shape[1])        # :44
    wce = torch.stack([weights[i] * yt[:, i] * torch.log(yh[:, i] + 1e-12) for i in range(num_classes)], dim=1)
    wce = -wce.sum(dim=1)                                     # :54
    return wce.mean() if reduction == "mean" else wce.sum()   # :55


def weighted_partial_ce_torch_ce(y_hat: Tensor, ys: Tensor, full: bool = False) -> Tensor:
    """architectures/segmentor/losses.py:26-61 with the constructor default manual=False and two classes: the masked
    y_hat (:31-32) is fed as LOGITS to nn.CrossEntropyLoss with the class index taken from ys[:,1:] (:40-45,58); the class
    weights of :34-38 and `reduction` are not used on this branch."""
    assert y_hat.shape[1] == ys.shape[1] == 2
    z = y_hat if full else y_hat * ys
    z = z.permute(0, 2, 3, 1).reshape(-1, 2)
    t = ys[:, 1:].permute(0, 2, 3, 1).reshape(-1).long()
    return torch.nn.functional.cross_entropy(z, t)


def weighted_partial_ce_bce(y_hat: Tensor, ys: Tensor, full: bool = False) -> Tensor:
    """architectures/segmentor/losses.py:48-49 (num_classes == 1, manual=True): nn.BCEWithLogitsLoss on the masked map."""
    assert y_hat.shape[1] == ys.shape[1] == 1
    z = y_hat if full else y_hat * ys
    return torch.nn.functional.binary_cross_entropy_with_logits(z.permute(0, 2, 3, 1).reshape(-1, 1), ys.permute(0, 2, 3, 1).reshape(-1, 1))


def dice_loss(inp: Tensor, target: Tensor, eps: float = 1e-12) -> Tensor:
    """architectures/segmentor/losses.py:70-74."""
    intersect = (inp * target).sum(dim=(1, 2, 3))
    cardinal = (inp + target).sum(dim=(1, 2, 3))
    return (-(2.0 * intersect / (cardinal + eps)) + 1.0).mean()


def nearest_resize(att: Tensor, size: Tuple[int, int]) -> Tensor:
    """kornia.geometry.transform.resize(..., interpolation='nearest') as used at losses.py:126:
    torch's legacy nearest, src = min(floor(dst * in/out), in-1)."""
    h, w = att.shape[-2:]
    H, W = size
    if (h, w) == (H, W):
        return att
    sy = torch.clamp(torch.floor(torch.arange(H, dtype=torch.float32) * (float(h) / H)).long(), max=h - 1)
    sx = torch.clamp(torch.floor(torch.arange(W, dtype=torch.float32) * (float(w) / W)).long(), max=w - 1)
    return att[:, :, sy][:, :, :, sx]


def interlayer_divergence(attentions: Sequence[Tensor], weights: Optional[list] = None,
                          stop_gradient: bool = False) -> Tensor:
    """architectures/segmentor/losses.py:111-147 (divergence='KLD', mode='mean')."""
    basis = attentions[0].detach() if stop_gradient else attentions[0]           # :114
    C = basis.shape[1]
    flat = lambda t: t.permute(0, 2, 3, 1).reshape(-1, C)
    log_basis = torch.log(flat(basis) + 1e-12)                                   # :115
    height, width = basis.shape[2], basis.shape[3]
    if weights is None:
        weights = [1 for _ in range(len(attentions[1:]))]                        # :118-119
    elif len(weights) != len(attentions[1:]):
        weights = weights[:len(attentions)]                                      # :121-123
    posterior = []
    for att, weight in zip(attentions[1:], weights):
        if weight == 0:
            continue                                                             # :125
        posterior.append(nearest_resize(att, (height, width)) * weight)          # :126
    post = torch.stack([flat(p) for p in posterior], dim=0)                      # :130
    m_log_prob = torch.log(post + 1e-12).sum(dim=0) / sum(weights)               # :135
    divergence = (flat(basis) * (log_basis - m_log_prob)).sum(dim=1).mean()      # :137-139
    return divergence


def interlayer_divergence_jsd(attentions: Sequence[Tensor], weights: Optional[list] = None,
                              stop_gradient: bool = False, eps: float = 1e-12) -> Tensor:
    """architectures/segmentor/losses.py:111-126,154-169 (divergence='JSD', mode='mean')."""
    basis = attentions[0].detach() if stop_gradient else attentions[0]           # :114
    C = basis.shape[1]
    flat = lambda t: t.permute(0, 2, 3, 1).reshape(-1, C)
    log_basis = torch.log(flat(basis) + 1e-12)                                   # :115
    height, width = basis.shape[2], basis.shape[3]
    if weights is None:
        weights = [1 for _ in range(len(attentions[1:]))]
    elif len(weights) != len(attentions[1:]):
        weights = weights[:len(attentions)]
    posterior = [nearest_resize(att, (height, width)) * weight
                 for att, weight in zip(attentions[1:], weights) if weight != 0]  # :124-126
    mean_q = torch.stack(posterior, dim=0).mean(dim=0)                           # :156-157
    mixture = 0.5 * (basis + mean_q)                                             # :158
    log_mixture = torch.log(flat(mixture) + eps)                                 # :159
    log_mean_q = flat(torch.log(mean_q + 1e-12))                                 # :160 (the 'mean' reduce has no reduced axis)
    kld_p = (0.5 * flat(basis) * (log_basis - log_mixture)).sum(dim=1).mean()    # :162-164
    kld_q = (0.5 * flat(mean_q) * (log_mean_q - log_mixture)).sum(dim=1).mean()  # :166-168
    return kld_p + kld_q


def ls_discriminator_loss(y_real: Tensor, y_fake: Tensor) -> Tensor:
    """architectures/discriminator/losses.py:11-14."""
    return 0.5 * torch.mean((y_real - 1) ** 2) + 0.5 * torch.mean((y_fake + 1) ** 2)


def ls_generator_loss(y_fake: Tensor) -> Tensor:
    """architectures/discriminator/losses.py:22-24."""
    return 0.5 * torch.mean((y_fake - 1) ** 2)


# =================================================================================================
# Network blocks (functional; `sd` maps reference state_dict keys -> tensors; `p` is the key prefix)
# =================================================================================================
class BNState:
    """Collects BatchNorm running-stat updates so tests can compare them with the CUDA path."""

    def __init__(self):
        self.updated: Dict[str, Tensor] = {}


def batch_norm(sd: Mapping[str, Tensor], p: str, x: Tensor, training: bool, st: Optional[BNState]) -> Tensor:
    """torch.nn.BatchNorm2d (momentum 0.1, eps 1e-5, biased var for normalisation, unbiased for running)."""
    rm = sd[p + "running_mean"].clone()
    rv = sd[p + "running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[p + "weight"], sd[p + "bias"], training, 0.1, 1e-5)
    if training and st is not None:
        st.updated[p + "running_mean"] = rm
        st.updated[p + "running_var"] = rv
        st.updated[p + "num_batches_tracked"] = sd[p + "num_batches_tracked"] + 1
    return y


def splat_conv(sd, p: str, x: Tensor, groups: int, training: bool, st, radix: int = 2) -> Tensor:
    """SplAtConv2d.forward, architectures/extra/resnest.py:97-138 (3x3, stride 1, padding 1)."""
    bias = sd.get(p + "conv.bias")
    x = F.conv2d(x, sd[p + "conv.weight"], bias, 1, 1, 1, groups * radix)        # :99
    x = batch_norm(sd, p + "bn0.", x, training, st)                              # :101
    x = F.relu(x)                                                                # :105
    batch, channel = x.shape[:2]
    splited = torch.split(x, channel // radix, dim=1)                            # :109
    gap = sum(splited)                                                           # :111
    gap = F.adaptive_avg_pool2d(gap, 1)                                          # :116
    gap = F.conv2d(gap, sd[p + "fc1.weight"], sd[p + "fc1.bias"], groups=groups)  # :118
    gap = batch_norm(sd, p + "bn1.", gap, training, st)                          # :121
    gap = F.relu(gap)
    channels = channel // radix
    atten = F.conv2d(gap, sd[p + "fc2.weight"], sd[p + "fc2.bias"], groups=groups).view(batch, radix, channels)  # :125
    atten = F.softmax(atten, dim=1).view(batch, -1, 1, 1)                        # :127
    atten = torch.split(atten, channel // radix, dim=1)                          # :133
    out = sum([a * s for a, s in zip(atten, splited)])                           # :135
    return out.contiguous()


def bottleneck(sd, p: str, x: Tensor, stride: int, has_down: bool, avd: bool, training: bool, st) -> Tensor:
    """Bottleneck.forward, resnest.py:234-267 with radix=2, cardinality=1, avd=True, avd_first=False."""
    residual = x
    out = F.conv2d(x, sd[p + "conv1.weight"])                                    # :237
    out = F.relu(batch_norm(sd, p + "bn1.", out, training, st))
    out = splat_conv(sd, p + "conv2.", out, 1, training, st)                     # :246
    if avd:
        out = F.avg_pool2d(out, 3, stride, padding=1)                            # :253-254 (count_include_pad=True)
    out = F.conv2d(out, sd[p + "conv3.weight"])
    out = batch_norm(sd, p + "bn3.", out, training, st)
    if has_down:                                                                 # resnest.py:380-394
        r = x
        if stride != 1:
            r = F.avg_pool2d(r, stride, stride, ceil_mode=True, count_include_pad=False)
        else:
            r = F.avg_pool2d(r, 1, 1, ceil_mode=True, count_include_pad=False)
        r = F.conv2d(r, sd[p + "downsample.1.weight"])
        residual = batch_norm(sd, p + "downsample.2.", r, training, st)
    return F.relu(out + residual)


def resnest_layer(sd, p: str, x: Tensor, blocks: int, stride: int, is_first: bool, training: bool, st) -> Tensor:
    """ResNet._make_layer, resnest.py:376-429.  avd is active when stride>1 or is_first (:187)."""
    x = bottleneck(sd, p + "0.", x, stride, True, stride > 1 or is_first, training, st)
    for i in range(1, blocks):
        x = bottleneck(sd, p + f"{i}.", x, 1, False, False, training, st)
    return x


def stem(sd, p: str, x: Tensor, training: bool, st) -> Tensor:
    """deep stem conv1 + bn1 + relu, resnest.py:326-339 (compose.py:40-44)."""
    x = F.conv2d(x, sd[p + "0.0.weight"], None, 2, 1)
    x = F.relu(batch_norm(sd, p + "0.1.", x, training, st))
    x = F.conv2d(x, sd[p + "0.3.weight"], None, 1, 1)
    x = F.relu(batch_norm(sd, p + "0.4.", x, training, st))
    x = F.conv2d(x, sd[p + "0.6.weight"], None, 1, 1)
    return F.relu(batch_norm(sd, p + "1.", x, training, st))


def decoder_block(sd, p: str, x: Tensor, training: bool, st) -> Tensor:
    """ResNestDecoder.forward, resnest.py:38-43."""
    residual = F.conv2d(x, sd[p + "downsample.0.weight"])
    residual = batch_norm(sd, p + "downsample.1.", residual, training, st)
    out = F.conv2d(x, sd[p + "conv.0.weight"], None, 1, 1)
    out = F.relu(batch_norm(sd, p + "conv.1.", out, training, st))
    out = F.relu(splat_conv(sd, p + "conv.3.", out, 2, training, st))
    return F.relu(residual + out)


def upsampling(sd, p: str, x: Tensor) -> Tensor:
    """Upsampling.forward, resnest.py:52-54 (ConvTranspose2d k=2 s=2 with bias)."""
    return F.conv_transpose2d(x, sd[p + "up.weight"], sd[p + "up.bias"], stride=2)


def attention_gate(sd, p: str, x: Tensor) -> Tuple[Tensor, Tensor]:
    """AdversarialAttentionGate.forward, segmentor/blocks.py:38-46."""
    y_hat = F.softmax(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"]), dim=1)
    mask = y_hat[:, 1:].sum(dim=1, keepdim=True)
    return x * mask, y_hat


def segmentor_forward(sd: Mapping[str, Tensor], x: Tensor, training: bool = True, gating_level: int = 4,
                      st: Optional[BNState] = None, p: str = ""):
    """ResnestUNet.forward, segmentor/compose.py:100-187 (encoder_gating=False).
    -> (attentions tuple full-res first, agg_map logits, x_4)."""
    x_0_0 = stem(sd, p + "encoder_0_1_2.", x, training, st)                      # :102
    x_0_1 = F.max_pool2d(x_0_0, 3, 2, 1)                                         # :103
    x_1 = resnest_layer(sd, p + "encoder_1.", x_0_1, 3, 1, False, training, st)  # :109
    x_2 = resnest_layer(sd, p + "encoder_2.", x_1, 4, 2, True, training, st)
    x_3 = resnest_layer(sd, p + "encoder_3.", x_2, 6, 2, True, training, st)
    down_padding = right_padding = False
    if x_3.shape[2] % 2 == 1:
        x_3 = F.pad(x_3, (0, 0, 0, 1)); down_padding = True                      # :125-127
    if x_3.shape[3] % 2 == 1:
        x_3 = F.pad(x_3, (0, 1, 0, 0)); right_padding = True                     # :128-130
    x_4 = resnest_layer(sd, p + "encoder_4.", x_3, 3, 2, True, training, st)     # :132
    attentions = []
    d_4 = torch.cat((x_3, upsampling(sd, p + "upsampling_4.", x_4)), dim=1)      # :140-141
    if down_padding:
        d_4 = d_4[:, :, :-1, :]
    if right_padding:
        d_4 = d_4[:, :, :, :-1]                                                  # :142-147
    d_4 = decoder_block(sd, p + "decoder_4.", d_4, training, st)
    if gating_level >= 4:
        d_4, y_4 = attention_gate(sd, p + "aag_4.", d_4); attentions.append(y_4)
    d_3 = torch.cat((x_2, upsampling(sd, p + "upsampling_3.", d_4)), dim=1)
    d_3 = decoder_block(sd, p + "decoder_3.", d_3, training, st)
    if gating_level >= 3:
        d_3, y_3 = attention_gate(sd, p + "aag_3.", d_3); attentions.append(y_3)
    d_2 = torch.cat((x_1, upsampling(sd, p + "upsampling_2.", d_3)), dim=1)
    d_2 = decoder_block(sd, p + "decoder_2.", d_2, training, st)
    if gating_level >= 2:
        d_2, y_2 = attention_gate(sd, p + "aag_2.", d_2); attentions.append(y_2)
    d_1 = torch.cat((x_0_0, upsampling(sd, p + "upsampling_1.", d_2)), dim=1)
    d_1 = decoder_block(sd, p + "decoder_1.", d_1, training, st)
    if gating_level >= 1:
        d_1, y_1 = attention_gate(sd, p + "aag_1.", d_1); attentions.append(y_1)
    d_0 = upsampling(sd, p + "upsampling_0.", d_1)                               # :175 (no skip)
    d_0 = decoder_block(sd, p + "decoder_0.", d_0, training, st)
    if gating_level >= 0:
        d_0, y_0 = attention_gate(sd, p + "aag_0.", d_0); attentions.append(y_0)
    agg_map = F.conv2d(d_0, sd[p + "fc.weight"], sd[p + "fc.bias"])              # :181
    attentions.reverse()                                                         # :183
    return tuple(attentions), agg_map, x_4


def parallel_head_forward(sd: Mapping[str, Tensor], x: Tensor, training: bool = True,
                          gating_level: Optional[int] = None, st: Optional[BNState] = None, p: str = ""):
    """ResnestUnetParallelHead.forward (segmentor/compose.py:291-350) when gating_level is None, else
    ResnestUnetParallelHeadAttentionGate.forward (:432-515; constructor default gating_leveL=3, level 4 gated only
    when gating_level > 3).  A second decoder pair (`*_c`) hangs off the encoder's x_1 and shares the x_0_0 skip.
    -> agg [2,B,C,H,W]   or   ((attentions), (attentions_c)), agg   with both tuples full-res first."""
    gated = gating_level is not None
    gl = gating_level if gated else -1
    x_0_0 = stem(sd, p + "encoder_0_1_2.", x, training, st)
    x_0_1 = F.max_pool2d(x_0_0, 3, 2, 1)
    x_1 = resnest_layer(sd, p + "encoder_1.", x_0_1, 3, 1, False, training, st)
    x_2 = resnest_layer(sd, p + "encoder_2.", x_1, 4, 2, True, training, st)
    x_3 = resnest_layer(sd, p + "encoder_3.", x_2, 6, 2, True, training, st)
    down_padding = right_padding = False
    if x_3.shape[2] % 2 == 1:
        x_3 = F.pad(x_3, (0, 0, 0, 1)); down_padding = True
    if x_3.shape[3] % 2 == 1:
        x_3 = F.pad(x_3, (0, 1, 0, 0)); right_padding = True
    x_4 = resnest_layer(sd, p + "encoder_4.", x_3, 3, 2, True, training, st)
    attentions, attentions_c = [], []

    def level(name: str, inp: Tensor, on: bool, sink: list) -> Tensor:
        d = decoder_block(sd, p + "decoder_" + name + ".", inp, training, st)
        if on:
            d, y = attention_gate(sd, p + "aag_" + name + ".", d)
            sink.append(y)
        return d

    d_4 = torch.cat((x_3, upsampling(sd, p + "upsampling_4.", x_4)), dim=1)
    if down_padding:
        d_4 = d_4[:, :, :-1, :]
    if right_padding:
        d_4 = d_4[:, :, :, :-1]
    d_4 = level("4", d_4, gated and gl > 3, attentions)
    d_3 = level("3", torch.cat((x_2, upsampling(sd, p + "upsampling_3.", d_4)), dim=1), gated and gl >= 3, attentions)
    d_2 = level("2", torch.cat((x_1, upsampling(sd, p + "upsampling_2.", d_3)), dim=1), gated and gl >= 2, attentions)
    d_1 = level("1", torch.cat((x_0_0, upsampling(sd, p + "upsampling_1.", d_2)), dim=1), gated and gl >= 1, attentions)
    d_0 = level("0", upsampling(sd, p + "upsampling_0.", d_1), gated and gl >= 0, attentions)
    # parallel branch: from the layer1 features, not from d_2
    d_1_c = level("1_c", torch.cat((x_0_0, upsampling(sd, p + "upsampling_1_c.", x_1)), dim=1), gated and gl >= 1, attentions_c)
    d_0_c = level("0_c", upsampling(sd, p + "upsampling_0_c.", d_1_c), gated and gl >= 0, attentions_c)
    agg = torch.stack([F.conv2d(d_0, sd[p + "fc.weight"], sd[p + "fc.bias"]),
                       F.conv2d(d_0_c, sd[p + "fc_c.weight"], sd[p + "fc_c.bias"])])
    if not gated:
        return agg
    attentions.reverse(); attentions_c.reverse()
    return (tuple(attentions), tuple(attentions_c)), agg


# =================================================================================================
# Discriminator
# =================================================================================================
def spectral_weight(sd, p: str, training: bool, updated: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """torch.nn.utils.spectral_norm (legacy hook), n_power_iterations=1, eps=1e-12, dim=0:
    one power iteration per training forward, u/v updated in place, sigma = u^T W v."""
    w = sd[p + "weight_orig"]
    u, v = sd[p + "weight_u"].clone(), sd[p + "weight_v"].clone()
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
        if updated is not None:
            updated[p + "weight_u"], updated[p + "weight_v"] = u, v
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def discriminator_forward(sd: Mapping[str, Tensor], ys: Sequence[Tensor], depth: int = 4, training: bool = True,
                          is_training_flag: bool = True, noise: Optional[Tensor] = None, instance_noise: bool = False,
                          flip: bool = False, p: str = "", updated: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """DiscriminatorBlock.forward, discriminator/blocks.py:114-130.
    `noise` ([H,W] plane) and `flip` replace the CPU RNG draws of InstanceNoise (:150) and LabelNoise (:166).
    `training` = nn.Module.training (drives the spectral-norm power iteration); `is_training_flag` = the
    constructor's is_training (drives whether noise is added, :151)."""
    s = ys[0]
    ci = 0
    if instance_noise:
        if is_training_flag and noise is not None:
            s = s + noise                                                        # :151
        s = torch.clip(s, 0, 1)                                                  # :152-153
        ci = 1
    s = F.conv2d(s, sd[p + f"stack_0.{ci}.weight"], sd[p + f"stack_0.{ci}.bias"], 2, 1)
    s = F.leaky_relu(s, 0.2)
    for i in range(depth):
        q = p + f"squeeze_dict.squeeze_{i}.0."
        s = torch.sigmoid(F.conv2d(s, sd[q + "weight"], sd[q + "bias"]))          # :121
        s = torch.cat((s, ys[i + 1]), dim=1)                                     # :122
        q = p + f"spectral_dict.spectral_{i}.0."
        s = torch.tanh(F.conv2d(s, spectral_weight(sd, q, training, updated), sd[q + "bias"], 2, 1))  # :123
    logits = F.conv2d(s, sd[p + "out.0.weight"], sd[p + "out.0.bias"]).flatten(1)  # :128
    if flip:
        logits = -1 * logits                                                     # :167-168
    return logits
