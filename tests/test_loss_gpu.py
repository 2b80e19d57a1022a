"""Parity of the fused CUDA loss kernels (through the reference-named modules and the C-ABI) with
(1) golden vectors produced by the real reference, (2) the CPU oracle on seeded inputs, (3) the
closed-form known answers of SURVEY.md §8c.  Tolerances: fp32 1e-4 relative, bf16 1e-2 relative
(BASELINE.json north_star)."""
import os

import pytest
import torch

from oracle import octave_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def L():
    from octave_b200 import losses
    return losses


def close(a, b, rtol=1e-4, atol=1e-6, msg=""):
    torch.testing.assert_close(a.detach().float().cpu(), b.detach().float().cpu(), rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def grad_close(a, b, rtol, msg=""):
    # gradients: relative to the largest magnitude of the reference gradient (element-wise rtol is
    # meaningless for entries that are ~0 by cancellation)
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = b.abs().max().clamp_min(1e-20)
    err = (a - b).abs().max() / scale
    assert err <= rtol, f"{msg}: max err / max|ref| = {err:.3e} > {rtol}"


@pytest.fixture(scope="module")
def cases():
    return torch.load(os.path.join(GOLD, "losses.pt"))


@pytest.mark.parametrize("name", ["pyr32", "pyr48x80", "generic"])
def test_losses_against_reference_golden_fp32(cases, name):
    c = cases[name]
    C = c["yhat"].shape[1]
    m = L()
    yh = c["yhat"].to(DEV).requires_grad_()
    wp = m.WeightedPartialCE(C, manual=True)
    l = wp(yh, c["ys"].to(DEV)); close(l, c["wpce"], msg="wpce")
    grad_close(torch.autograd.grad(l, yh)[0], c["wpce_g"], 1e-4, "wpce grad")
    close(wp(yh, c["ys"].to(DEV), reduction="sum"), c["wpce_sum"], msg="wpce sum")
    l = wp(yh, c["ys"].to(DEV), full=True); close(l, c["wpce_full"], msg="wpce full")
    grad_close(torch.autograd.grad(l, yh)[0], c["wpce_full_g"], 1e-4, "wpce full grad")
    ysb = c["ys"].to(DEV)
    close(wp(yh, ysb, ignore_bg=True), c["wpce_ignore_bg"], msg="ignore_bg")
    assert torch.equal(ysb.cpu(), c["ys_after_ignore_bg"]), "ignore_bg must zero ys[:,0] in the caller's tensor"
    lg = c["logits"].to(DEV).requires_grad_()
    l = wp(lg, c["ys"].to(DEV), from_logits=True); close(l, c["wpce"], msg="wpce from logits")
    grad_close(torch.autograd.grad(l, lg)[0], c["wpce_logits_g"], 1e-4, "wpce logits grad")
    l = m.DiceLoss()(yh, c["full"].to(DEV)); close(l, c["dice"], msg="dice")
    grad_close(torch.autograd.grad(l, yh)[0], c["dice_g"], 1e-4, "dice grad")
    att = [a.to(DEV).requires_grad_() for a in c["att"]]
    l = m.InterlayerDivergence()(att); close(l, c["kld"], msg="kld")
    for k, (g, gg) in enumerate(zip(torch.autograd.grad(l, att), c["kld_g"])):
        grad_close(g, gg, 1e-4, f"kld grad level {k}")
    l = m.InterlayerDivergence()(att, weights=c["kld_weights"]); close(l, c["kld_w"], msg="kld weights")
    for k, (g, gg) in enumerate(zip(torch.autograd.grad(l, att), c["kld_w_g"])):
        if gg is None:
            assert float(g.abs().max()) == 0.0
        else:
            grad_close(g, gg, 1e-4, f"kld weighted grad level {k}")
    l = m.InterlayerDivergence(stop_gradient=True)(att)
    gs = torch.autograd.grad(l, att)
    assert float(gs[0].abs().max()) == 0.0
    for k in range(1, len(att)):
        grad_close(gs[k], c["kld_stop_g"][k], 1e-4, f"kld stopgrad level {k}")
    df = c["d_fake"].to(DEV).requires_grad_(); dr = c["d_real"].to(DEV).requires_grad_()
    l = m.LSGeneratorLoss()(df); close(l, c["lsg"], msg="lsg")
    grad_close(torch.autograd.grad(l, df)[0], c["lsg_g"], 1e-5, "lsg grad")
    l = m.LSDiscriminatorialLoss()(dr, df); close(l, c["lsd"], msg="lsd")
    g = torch.autograd.grad(l, (dr, df))
    grad_close(g[0], c["lsd_g"][0], 1e-5, "lsd grad real"); grad_close(g[1], c["lsd_g"][1], 1e-5, "lsd grad fake")


def test_known_answers():
    m = L()
    ys = torch.zeros(1, 2, 2, 2, device=DEV); ys[0, 0, 0, 0] = 1; ys[0, 1, 0, 1] = 1; ys[0, 1, 1, 0] = 1
    yh = torch.full((1, 2, 2, 2), 0.5, device=DEV, requires_grad=True)
    wp = m.WeightedPartialCE(2, manual=True)
    l = wp(yh, ys.clone()); assert abs(l.item() - 1.0397208) < 1e-5
    g, = torch.autograd.grad(l, yh)
    close(g.flatten(), torch.tensor([-1.5, 0, 0, 0, 0, -0.75, -0.75, 0]), rtol=1e-5)
    assert abs(wp(yh, ys.clone(), reduction='sum').item() - 4.1588831) < 1e-4
    y2 = ys.clone(); assert abs(wp(yh, y2, ignore_bg=True).item() - 0.3465736) < 1e-5; assert float(y2[:, 0].abs().sum()) == 0
    assert wp(yh, torch.zeros_like(ys)).item() == 0.0  # no scribbles: 0, not NaN
    p = torch.full((2, 2, 2, 2), 0.5, device=DEV); t = torch.zeros(2, 2, 2, 2, device=DEV); t[:, 1] = 1
    assert abs(m.DiceLoss()(p, t).item() - 0.5) < 1e-6
    b = torch.zeros(1, 2, 4, 4, device=DEV); b[:, 0] = .8; b[:, 1] = .2
    q1 = torch.full((1, 2, 2, 2), .5, device=DEV); q2 = torch.zeros(1, 2, 1, 1, device=DEV); q2[:, 0] = .25; q2[:, 1] = .75
    assert abs(m.InterlayerDivergence()([b, q1, q2]).item() - 0.4294572) < 1e-5
    assert abs(m.InterlayerDivergence()([b, q1, q2], weights=[2, 0]).item() + 0.5004024) < 1e-5
    r = torch.tensor([[.5], [1.5]], device=DEV); f = torch.tensor([[-.5], [0.]], device=DEV)
    assert abs(m.LSDiscriminatorialLoss()(r, f).item() - 0.4375) < 1e-6
    assert abs(m.LSGeneratorLoss()(f).item() - 0.8125) < 1e-6


def test_error_behaviour():
    m = L()
    with pytest.raises(AssertionError, match="Number of class mismatch"):
        m.WeightedPartialCE(2, manual=True)(torch.rand(1, 2, 16, 16, device=DEV), torch.rand(1, 3, 16, 16, device=DEV))
    with pytest.raises(NotImplementedError):
        m.InterlayerDivergence(mode='sum')([torch.rand(1, 2, 16, 16, device=DEV)] * 2)
    with pytest.raises(NotImplementedError):
        m.InterlayerDivergence(divergence='XYZ')([torch.rand(1, 2, 16, 16, device=DEV)] * 2)
    bad = torch.full((1, 2, 16, 16), float('nan'), device=DEV)
    with pytest.raises(Exception, match="Divergence is NaN"):
        m.InterlayerDivergence()([bad, bad[:, :, ::2, ::2].contiguous()])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.DiceLoss()(torch.rand(1, 2, 16, 16), torch.rand(1, 2, 16, 16))


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 304, 304), (3, 400, 400), (1, 16, 16), (2, 48, 1024)])
def test_fused_g_step_loss_vs_oracle(dtype, rtol, shape):
    """FusedSegmentorLoss (all G-step terms in one launch pair) vs the oracle on synthetic OCTA inputs."""
    B, H, W = shape
    m = L()
    g = torch.Generator().manual_seed(H * 7 + W)
    logits = (2 * torch.randn(B, 2, H, W, generator=g)).to(dtype)
    _, ys, _ = synth.octa_batch(B, H, W, seed=3)
    att = [a.to(dtype) for a in synth.prob_maps(B, 2, H, W, 5, seed=9)]
    d_fake = torch.randn(B, 1, generator=g)
    # oracle on the same (dtype-rounded) values, in fp64
    lo = logits.double().requires_grad_(); ao = [a.double().requires_grad_() for a in att]; fo = d_fake.double().requires_grad_()
    w_o = O.weighted_partial_ce(torch.softmax(lo, 1), ys.double(), 2)
    k_o = O.interlayer_divergence(ao)
    g_o = O.ls_generator_loss(fo)
    tot_o = w_o + 0.5 * k_o + 0.25 * g_o
    grads_o = torch.autograd.grad(tot_o, [lo, *ao, fo])
    lc = logits.to(DEV).requires_grad_(); ac = [a.to(DEV).requires_grad_() for a in att]; fc = d_fake.to(DEV).requires_grad_()
    res = m.FusedSegmentorLoss()(lc, ys.to(DEV).to(dtype), ac, fc)
    close(res['supervised'], w_o, rtol=rtol, msg="wpce"); close(res['divergence'], k_o, rtol=rtol, msg="kld")
    close(res['generator'], g_o, rtol=1e-5, msg="lsg")
    tot = res['supervised'] + 0.5 * res['divergence'] + 0.25 * res['generator']
    grads = torch.autograd.grad(tot, [lc, *ac, fc])
    for i, (a, b) in enumerate(zip(grads, grads_o)):
        grad_close(a, b, rtol, f"grad[{i}]")
    assert float(res['nan_flag']) == 0.0


def test_generic_path_vs_oracle():
    """C=3, map sizes that are not a 2^k pyramid: exercises the generic kernel (atomics for coarse grads)."""
    m = L()
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 3, 37, 51
    yh = torch.softmax(torch.randn(B, C, H, W, generator=g), 1)
    lab = torch.randint(0, C + 2, (B, H, W), generator=g)
    ys = torch.stack([(lab == c).float() for c in range(C)], 1)
    att = [torch.softmax(torch.randn(B, C, h, w, generator=g), 1) for h, w in [(H, W), (19, 25), (9, 13), (4, 7)]]
    yo = yh.double().requires_grad_(); ao = [a.double().requires_grad_() for a in att]
    tot_o = O.weighted_partial_ce(yo, ys.double(), C) + O.dice_loss(yo, ys.double()) + O.interlayer_divergence(ao, [1, 0.5, 2])
    go = torch.autograd.grad(tot_o, [yo, *ao])
    yc = yh.to(DEV).requires_grad_(); ac = [a.to(DEV).requires_grad_() for a in att]
    tot = m.WeightedPartialCE(C, manual=True)(yc, ys.to(DEV)) + m.DiceLoss()(yc, ys.to(DEV)) + m.InterlayerDivergence()(ac, [1, 0.5, 2])
    close(tot, tot_o, rtol=1e-4)
    for i, (a, b) in enumerate(zip(torch.autograd.grad(tot, [yc, *ac]), go)):
        grad_close(a, b, 1e-4, f"grad[{i}]")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("weights,stop", [(None, False), ([1, 0.5, 2], False), ([1, 0, 2], True)])
def test_interlayer_divergence_jsd_vs_oracle(weights, stop, dtype):
    """divergence='JSD' (reference losses.py:154-169) on the generic kernel: value, gradients, stop_gradient, weights."""
    m = L()
    g = torch.Generator().manual_seed(11)
    B, C = 2, 2
    sizes = [(48, 64), (24, 32), (12, 16), (6, 8)]
    att = [torch.softmax(torch.randn(B, C, h, w, generator=g), 1).to(dtype).float() for h, w in sizes]
    ao = [a.double().requires_grad_() for a in att]
    lo = O.interlayer_divergence_jsd(ao, weights, stop_gradient=stop)
    go = torch.autograd.grad(lo, ao, allow_unused=True)
    ac = [a.to(DEV).to(dtype).requires_grad_() for a in att]
    lc = m.InterlayerDivergence(divergence='JSD', stop_gradient=stop)(ac, weights)
    close(lc, lo, rtol=1e-4)
    gc = torch.autograd.grad(lc, ac, allow_unused=True)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    for i, (a, b) in enumerate(zip(gc, go)):
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.float().abs().max()) == 0.0, f"grad[{i}] should be zero"
        else:
            grad_close(a.float(), b, tol, f"grad[{i}]")
    # known answer
    b0 = torch.zeros(1, 2, 4, 4); b0[:, 0] = .8; b0[:, 1] = .2
    q1 = torch.full((1, 2, 2, 2), .5); q2 = torch.zeros(1, 2, 1, 1); q2[:, 0] = .25; q2[:, 1] = .75
    assert abs(m.InterlayerDivergence(divergence='JSD')([b0.to(DEV), q1.to(DEV), q2.to(DEV)]).item() - 0.0967728) < 1e-5


def test_full_size_properties():
    """BASELINE config sizes (c5: B=8, 1024^2): size-independent properties instead of a CPU oracle run."""
    m = L()
    B, H, W = 8, 1024, 1024
    g = torch.Generator(device=DEV).manual_seed(0)
    logits = torch.randn(B, 2, H, W, device=DEV, generator=g)
    p = torch.softmax(logits, 1)
    # KLD of a pyramid that is the exact nearest-downsample of constant maps is zero; identical maps -> 0
    const = torch.zeros(B, 2, H, W, device=DEV); const[:, 0] = 0.3; const[:, 1] = 0.7
    pyr = [const[:, :, :: 2 ** k, :: 2 ** k].contiguous() for k in range(5)]
    assert abs(m.InterlayerDivergence()(pyr).item()) < 1e-6
    # Dice(p, p-hard) in [0,1]; Dice(t,t)=0 for one-hot t
    t = torch.nn.functional.one_hot(p.argmax(1), 2).permute(0, 3, 1, 2).float().contiguous()
    assert abs(m.DiceLoss()(t, t).item()) < 1e-6
    # linearity of WPCE in the mask for one class: loss(sum) with reduction='sum' is additive over disjoint scribble sets
    # when class weights are equal (balanced counts): compare fused-from-logits with module-on-probabilities
    ys = torch.zeros(B, 2, H, W, device=DEV); ys[:, 0, ::7, ::5] = 1; ys[:, 1, 3::7, 2::5] = 1
    a = m.WeightedPartialCE(2, manual=True)(p, ys)
    b = m.WeightedPartialCE(2, manual=True)(logits, ys, from_logits=True)
    assert abs(a.item() - b.item()) <= 1e-4 * abs(a.item())
    # gradient of the sum of probabilities-path and logits-path agree through softmax's Jacobian
    lg = logits.clone().requires_grad_()
    l1 = m.WeightedPartialCE(2, manual=True)(lg, ys, from_logits=True)
    g1, = torch.autograd.grad(l1, lg)
    assert float(g1.sum(dim=1).abs().max()) < 1e-6  # softmax-Jacobian rows sum to zero


# ---- single-pass G-step objective (octave_loss_fused): values and gradients against the CPU oracle ------------------------
def _gstep_inputs(B, H, W, dtype, seed=0, label_frac=0.05):
    g = torch.Generator().manual_seed(seed)
    agg = torch.randn(B, 2, H, W, generator=g)
    lab = torch.rand(B, H, W, generator=g)
    ys = torch.stack([(lab < label_frac).float(), ((lab >= label_frac) & (lab < 2 * label_frac)).float()], 1)
    att = [torch.softmax(2.0 * torch.randn(B, 2, H >> k, W >> k, generator=g), 1) for k in range(5)]
    fake = torch.randn(B, 1, generator=g)
    rd = lambda t: t.to(dtype).float()          # the oracle sees exactly the values the kernel reads
    return rd(agg), rd(ys), [rd(a) for a in att], fake


def _oracle_total(agg, ys, att, fake, lam):
    agg = agg.clone().requires_grad_(); att = [a.clone().requires_grad_() for a in att]; fake = fake.clone().requires_grad_()
    sup = O.weighted_partial_ce(torch.softmax(agg, 1), ys, 2)
    kld = O.interlayer_divergence(att)
    gen = O.ls_generator_loss(fake)
    tot = lam[0] * sup + lam[1] * kld + lam[2] * gen
    gs = torch.autograd.grad(tot, [agg, fake, *att])
    return (sup, kld, gen, tot), gs


@pytest.mark.parametrize("dtype,B,H,W", [(torch.float32, 2, 304, 304), (torch.float32, 3, 48, 80), (torch.bfloat16, 2, 304, 304),
                                         (torch.bfloat16, 4, 400, 400)])
def test_fused_single_pass_matches_oracle(dtype, B, H, W):
    m = L()
    lam = (1.0, 0.1, 0.25)
    agg, ys, att, fake = _gstep_inputs(B, H, W, dtype)
    (sup, kld, gen, tot), gref = _oracle_total(agg, ys, att, fake, lam)
    a_d = agg.to(DEV, dtype).requires_grad_(); t_d = [a.to(DEV, dtype).requires_grad_() for a in att]
    f_d = fake.to(DEV).requires_grad_()
    res = m.FusedSegmentorLoss().total(a_d, ys.to(DEV, dtype), t_d, f_d, *lam)
    rt = 1e-4 if dtype == torch.float32 else 1e-2
    close(res['supervised'], sup, rtol=rt, msg="fused wpce")
    close(res['divergence'], kld, rtol=rt, msg="fused kld")
    close(res['generator'], gen, rtol=1e-4, msg="fused lsg")
    close(res['total'], tot, rtol=rt, msg="fused total")
    assert float(res['nan_flag']) == 0.0
    (3.0 * res['total']).backward()             # a non-unit upstream gradient exercises octave_loss_scale_grads
    got = [a_d.grad, f_d.grad, *[t.grad for t in t_d]]
    for k, (g, gg) in enumerate(zip(got, gref)):
        grad_close(g, 3.0 * gg, rt, f"fused gradient {k}")


def test_fused_single_pass_equals_two_pass_and_is_deterministic():
    """Same arithmetic as the statistics pass + gradient pass (fp32: values to 1e-5, gradients to 1e-5 of their maximum),
    and bit-identical results from two runs (fixed-order reductions)."""
    m = L()
    agg, ys, att, fake = _gstep_inputs(4, 160, 208, torch.float32, seed=3)
    fl = m.FusedSegmentorLoss()
    outs = []
    for _ in range(2):
        a_d = agg.to(DEV).requires_grad_(); t_d = [a.to(DEV).requires_grad_() for a in att]; f_d = fake.to(DEV).requires_grad_()
        r = fl.total(a_d, ys.to(DEV), t_d, f_d, 1.0, 0.1, 0.1)
        r['total'].backward()
        outs.append((r['total'].detach().clone(), a_d.grad.clone(), [t.grad.clone() for t in t_d], f_d.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][3], outs[1][3])
    for x, y in zip(outs[0][2], outs[1][2]):
        assert torch.equal(x, y)
    a_d = agg.to(DEV).requires_grad_(); t_d = [a.to(DEV).requires_grad_() for a in att]; f_d = fake.to(DEV).requires_grad_()
    r2 = fl(a_d, ys.to(DEV), t_d, f_d)
    tot2 = r2['supervised'] + 0.1 * r2['divergence'] + 0.1 * r2['generator']
    tot2.backward()
    close(outs[0][0], tot2, rtol=1e-5, msg="single pass vs two pass total")
    grad_close(outs[0][1], a_d.grad, 1e-5, "agg gradient")
    grad_close(outs[0][3], f_d.grad, 1e-5, "critic-logit gradient")
    for k, (x, t) in enumerate(zip(outs[0][2], t_d)):
        grad_close(x, t.grad, 1e-5, f"attention gradient {k}")


def test_fused_single_pass_falls_back_for_dice():
    m = L()
    agg, ys, att, fake = _gstep_inputs(2, 64, 64, torch.float32, seed=5)
    full = torch.stack([1 - (att[0][:, 1] > 0.5).float(), (att[0][:, 1] > 0.5).float()], 1)
    a_d = agg.to(DEV).requires_grad_(); t_d = [a.to(DEV).requires_grad_() for a in att]
    res = m.FusedSegmentorLoss(weakly_supervise=False).total(a_d, full.to(DEV), t_d, None, 1.0, 0.5, 0.0)
    exp = O.dice_loss(torch.softmax(agg, 1), full) + 0.5 * O.interlayer_divergence(att)
    close(res['total'], exp, rtol=1e-4, msg="dice fallback total")
    res['total'].backward()
    assert a_d.grad is not None and torch.isfinite(a_d.grad).all()


def test_wpce_default_manual_false_and_single_class_branches():
    """WeightedPartialCE(num_classes) with the constructor default manual=False (nn.CrossEntropyLoss branch, two classes) and
    num_classes == 1 (nn.BCEWithLogitsLoss): values and gradients against the oracle restatement pinned to the reference."""
    m = L()
    g = torch.Generator().manual_seed(3)
    ys = (torch.rand(3, 2, 40, 56, generator=g) < 0.2).float(); ys[:, 0] *= 1 - ys[:, 1]
    for full in (False, True):
        a = torch.randn(3, 2, 40, 56, generator=g)
        ao = a.clone().requires_grad_(); lo = O.weighted_partial_ce_torch_ce(ao, ys, full=full)
        ad = a.to(DEV).requires_grad_(); ld = m.WeightedPartialCE(2)(ad, ys.to(DEV), full=full)
        close(ld, lo, msg="manual=False")
        grad_close(torch.autograd.grad(2.0 * ld, ad)[0], 2.0 * torch.autograd.grad(lo, ao)[0], 1e-4, "manual=False grad")
        a1 = torch.randn(3, 1, 40, 56, generator=g); y1 = ys[:, 1:].contiguous()
        ao = a1.clone().requires_grad_(); lo = O.weighted_partial_ce_bce(ao, y1, full=full)
        ad = a1.to(DEV).requires_grad_(); ld = m.WeightedPartialCE(1, manual=True)(ad, y1.to(DEV), full=full)
        close(ld, lo, msg="num_classes=1")
        grad_close(torch.autograd.grad(ld, ad)[0], torch.autograd.grad(lo, ao)[0], 1e-4, "num_classes=1 grad")
    ysb = ys.to(DEV)
    m.WeightedPartialCE(2)(torch.randn(3, 2, 40, 56, device=DEV), ysb, ignore_bg=True)
    assert float(ysb[:, 0].abs().max()) == 0.0            # in-place side effect of ignore_bg, as in the reference
    with pytest.raises(ValueError):
        m.WeightedPartialCE(3)(torch.rand(2, 3, 8, 8, device=DEV), torch.zeros(2, 3, 8, 8, device=DEV))
