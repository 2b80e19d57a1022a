"""Parity of the CUDA path with the oracle AT THE BASELINE CONFIGURATIONS (BASELINE.json configs[0], configs[1] shape):
  c1  batch 2, 304x304  (the reference's own CPU-runnable case)
  c2  400x400, batch 8  (configs[1]'s shape; B >= 8 so that SplAtConv2d.bn1, extra/resnest.py:120-122, which normalises a
                         [B, C] tensor over the batch, is conditioned)
north_star bars: per-loss values within 1e-4 (fp32 mode) / 1e-2 (bf16 mode) of the reference through the whole network,
bit-exact argmax vessel masks in fp32 mode.

Gradients.  In train mode (batch statistics) the parameter gradients of this network at random init are ill-conditioned:
the fp32 oracle itself deviates from its fp64 run by 0.6-1.8 % per encoder stage, and ANY bf16 evaluation loses the
encoder directions — measured here with stock torch ops under bf16 autocast (cuDNN) on the same inputs: whole-net cosine
0.931 vs 0.935 for these kernels (B=8, 400x400; tools/parity_diag.py, profiles/parity_r02.log).  The tests therefore bound
 (a) fp32 mode: every module within max(2e-2, 6x the fp32-oracle noise floor) of the fp64 oracle, cosine >= 0.999;
 (b) bf16 mode, train: the output-side modules (fc, gates, decoder_0/1) within 5e-2 (measured 1e-3 .. 3e-2), and the whole-net error not larger
     than 1.1x that of the reference arithmetic under torch's own bf16 autocast;
 (c) bf16 mode, eval-mode BatchNorm (well conditioned): whole-net cosine >= 0.999, norm ratio within 2 %."""
import pytest
import torch

from oracle import octave_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

ORDER = ["fc", "aag_0", "decoder_0", "upsampling_0", "aag_1", "decoder_1", "upsampling_1", "aag_2", "decoder_2", "upsampling_2", "aag_3",
         "decoder_3", "upsampling_3", "aag_4", "decoder_4", "upsampling_4", "encoder_4", "encoder_3", "encoder_2", "encoder_1", "encoder_0_1_2"]


def l2err(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _build(mode, seed=0, training=True):
    from octave_b200 import config, network
    config.set_compute_dtype(mode)
    torch.manual_seed(seed)
    net = network.ResnestUNet(2, False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.cuda().train(training), sd


def _oracle(sd, x, ys, training=True, device="cpu", autocast=False):
    sdr = {k: (v.clone().to(device).requires_grad_() if v.is_floating_point() and "running" not in k else v.clone().to(device))
           for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        att, agg, _ = O.segmentor_forward(sdr, x.to(device), training=training, st=O.BNState())
    att = [a.float() for a in att]
    wp = O.weighted_partial_ce(torch.softmax(agg.float(), 1), ys.to(device), 2)
    kl = O.interlayer_divergence(att)
    names = [k for k, v in sdr.items() if v.requires_grad and not k.startswith("linear_head_")]
    gs = torch.autograd.grad(wp + 0.1 * kl, [sdr[k] for k in names], allow_unused=True)
    return att, agg.float(), wp, kl, {k: g for k, g in zip(names, gs) if g is not None}


def _cuda(net, x, ys):
    from octave_b200 import losses
    att, agg, _ = net(x.cuda())
    res = losses.FusedSegmentorLoss().total(agg, ys.cuda(), att, None, 1.0, 0.1, 0.0)
    res['total'].backward()
    return att, agg, res['supervised'].detach(), res['divergence'].detach(), {k: p.grad for k, p in net.named_parameters() if p.grad is not None}


def _skip(k, training):
    # identically-zero gradients in exact arithmetic: biases in front of a train-mode BatchNorm (rounding noise in any implementation)
    return training and (k.endswith(("fc1.bias",)) or (".conv" in k and k.endswith(".bias") and "fc2" not in k))


def _module_errors(ga, gb, training=True):
    out = {}
    for m in ORDER:
        num = da = db = dd = 0.0
        for k, b in gb.items():
            if not k.startswith(m + ".") or k not in ga or _skip(k, training):
                continue
            a = ga[k].detach().float().cpu().double().flatten(); b = b.detach().cpu().double().flatten()
            num += float(a @ b); da += float(a @ a); db += float(b @ b); dd += float((a - b) @ (a - b))
        if db > 0:
            out[m] = ((dd / db) ** 0.5, num / (da ** 0.5 * db ** 0.5 + 1e-300))
    return out


def _whole(ga, gb, training=True):
    num = da = db = dd = 0.0
    for k, b in gb.items():
        if k not in ga or _skip(k, training):
            continue
        a = ga[k].detach().float().cpu().double().flatten(); b = b.detach().cpu().double().flatten()
        num += float(a @ b); da += float(a @ a); db += float(b @ b); dd += float((a - b) @ (a - b))
    return num / (da ** 0.5 * db ** 0.5), (da / db) ** 0.5, (dd / db) ** 0.5


def rel(a, b):
    return abs(float(a) - float(b)) / abs(float(b))


# ---------------------------------------------------------------------------------------------------------------------
def test_c1_fp32_values_masks_and_gradients():
    """configs[0]: batch 2, 304x304, fp32 mode, train-mode BatchNorm."""
    net, sd = _build("fp32")
    x, ys, _ = synth.octa_batch(2, 304, 304, seed=11)
    att_o, agg_o, wp_o, kl_o, g_o = _oracle(sd, x, ys)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    _, _, _, _, g_d = _oracle(sd64, x.double(), ys.double())
    att, agg, wp, kl, g = _cuda(net, x, ys)
    assert l2err(agg, agg_o) < 1e-4
    for a, b in zip(att, att_o):
        assert l2err(a, b) < 1e-4
    assert torch.equal(agg.argmax(1).cpu(), agg_o.argmax(1)), "fp32 argmax vessel mask must be bit-exact"
    assert rel(wp, wp_o) <= 1e-4 and rel(kl, kl_o) <= 1e-4, (float(wp), float(wp_o), float(kl), float(kl_o))
    floor, mine = _module_errors(g_o, g_d), _module_errors(g, g_d)
    for m, (e, c) in mine.items():
        assert e <= max(2e-2, 6 * floor[m][0]), f"{m}: gradient error {e:.3e} vs fp64 oracle (fp32-oracle floor {floor[m][0]:.3e})"
    cos, ratio, _ = _whole(g, g_d)
    assert cos >= 0.999 and abs(ratio - 1) < 1e-2, (cos, ratio)


def test_c2_shape_fp32_masks_bit_exact():
    """400x400 (the 25 -> 26 -> 25 pad / crop of compose.py:125-147), fp32 mode: values 1e-4, bit-exact masks."""
    net, sd = _build("fp32", seed=1)
    x, ys, _ = synth.octa_batch(2, 400, 400, seed=12)
    with torch.no_grad():
        att_o, agg_o, _ = O.segmentor_forward(sd, x, training=True, st=O.BNState())
        wp_o = O.weighted_partial_ce(torch.softmax(agg_o, 1), ys, 2); kl_o = O.interlayer_divergence(att_o)
    att, agg, wp, kl, _ = _cuda(net, x, ys)
    assert l2err(agg, agg_o) < 1e-4
    assert torch.equal(agg.argmax(1).cpu(), agg_o.argmax(1))
    assert rel(wp, wp_o) <= 1e-4 and rel(kl, kl_o) <= 1e-4


@pytest.mark.parametrize("B,H", [(2, 304), (8, 400)])
def test_bf16_losses_and_gradients_train(B, H):
    """bf16 mode (the benchmarked path), train-mode BatchNorm, at c1 and at c2's shape."""
    net, sd = _build("bf16", seed=B)
    x, ys, _ = synth.octa_batch(B, H, H, seed=13 + B)
    att_o, agg_o, wp_o, kl_o, g_o = _oracle(sd, x, ys)
    att, agg, wp, kl, g = _cuda(net, x, ys)
    assert rel(wp, wp_o) <= 1e-2 and rel(kl, kl_o) <= 1e-2, (float(wp), float(wp_o), float(kl), float(kl_o))
    assert l2err(agg, agg_o) <= 3e-2, l2err(agg, agg_o)
    for a, b in zip(att[:3], att_o[:3]):
        assert l2err(a, b) <= 2e-2
    mine = _module_errors(g, g_o)
    for m in ("fc", "aag_0", "decoder_0", "upsampling_0", "aag_1", "decoder_1"):
        assert mine[m][0] <= 5e-2 and mine[m][1] >= 0.998, (m, mine[m])
    # calibration: the reference arithmetic itself in bf16 (stock torch CUDA ops under autocast) against the same fp32 oracle
    _, _, wp_a, kl_a, g_a = _oracle(sd, x, ys, device="cuda", autocast=True)
    cos, ratio, err = _whole(g, g_o)
    cos_a, ratio_a, err_a = _whole(g_a, g_o)
    print(f"B={B} {H}x{H} bf16 train: whole-net gradient error {err:.3f} (cos {cos:.4f}); torch bf16 autocast {err_a:.3f} (cos {cos_a:.4f}); "
          f"loss errors wpce {rel(wp, wp_o):.2e} kld {rel(kl, kl_o):.2e} (autocast {rel(wp_a, wp_o):.2e} {rel(kl_a, kl_o):.2e})")
    if B >= 8:      # at B = 2 SplAtConv2d.bn1 normalises over two samples: both bf16 evaluations are noise there (error > 1)
        assert err <= 1.1 * err_a + 1e-2, (err, err_a)
        assert abs(ratio - 1) < 5e-2


def test_bf16_gradients_eval_mode_cosine():
    """bf16 mode with BatchNorm in eval mode (no batch statistics => well conditioned): whole-net cosine >= 0.999."""
    net, _ = _build("bf16", seed=2, training=False)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(1 + 0.2 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x, ys, _ = synth.octa_batch(8, 304, 304, seed=4)
    att_o, agg_o, wp_o, kl_o, g_o = _oracle(sd, x, ys, training=False)
    att, agg, wp, kl, gg = _cuda(net, x, ys)
    # with eval-mode statistics the five attention maps nearly coincide: the divergence is ~3e-3, so an absolute floor applies
    assert rel(wp, wp_o) <= 1e-2 and abs(float(kl) - float(kl_o)) <= 1e-2 * abs(float(kl_o)) + 3e-4, (float(kl), float(kl_o))
    cos, ratio, err = _whole(gg, g_o, training=False)
    print(f"bf16 eval-mode whole-net gradient: cosine {cos:.6f}, norm ratio {ratio:.4f}, relative L2 {err:.4f}")
    assert cos >= 0.999 and abs(ratio - 1) < 2e-2, (cos, ratio)


def test_train_step_forward_is_bit_reproducible():
    """Two runs of the same G-step give bit-identical outputs and loss values (fixed-order global average pool and loss
    reductions, fp64 BatchNorm statistics).  Gradients: deterministic mode removes the fp32 atomics of the split-K weight
    gradients and the K-split linears; the head / narrow-3x3 weight gradients still meet through fp32 atomics, so parameter
    gradients are reproducible only to rounding (reported, not asserted)."""
    from octave_b200 import config
    outs = []
    try:
        config.set_deterministic(True)
        for _ in range(2):
            net, _ = _build("bf16", seed=3)
            x, ys, _ = synth.octa_batch(4, 160, 160, seed=21)
            _, agg, wp, kl, g = _cuda(net, x, ys)
            outs.append((wp.clone(), kl.clone(), agg.clone(), {k: v.clone() for k, v in g.items()}))
    finally:
        config.set_deterministic(False)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    bad = [k for k in outs[0][3] if not torch.equal(outs[0][3][k], outs[1][3][k])]
    worst = max((float((outs[0][3][k] - outs[1][3][k]).abs().max() / outs[0][3][k].abs().max().clamp_min(1e-30)) for k in bad), default=0.0)
    print(f"deterministic mode: {len(bad)} of {len(outs[0][3])} gradients differ between two runs (largest relative difference {worst:.2e})")
