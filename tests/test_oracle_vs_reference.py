"""Pins the oracle (oracle/octave_oracle.py) against the real reference imported from /root/reference.
CPU only; skipped where the reference is not mounted (the GPU box)."""
import pytest
import torch

from oracle import octave_oracle as O
from oracle import refload
from tests import synth

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return refload.load()


def test_wpce_default_and_single_class_branches_match_reference(ref):
    """WeightedPartialCE(manual=False) (the constructor default: nn.CrossEntropyLoss branch) and num_classes == 1
    (nn.BCEWithLogitsLoss), losses.py:40-49,56-59 — values and gradients."""
    g = torch.Generator().manual_seed(3)
    ys = (torch.rand(3, 2, 12, 20, generator=g) < 0.2).float(); ys[:, 0] *= 1 - ys[:, 1]
    for full in (False, True):
        a = torch.randn(3, 2, 12, 20, generator=g).requires_grad_(); b = a.detach().clone().requires_grad_()
        lr = ref.WeightedPartialCE(2)(a, ys.clone(), full=full); lo = O.weighted_partial_ce_torch_ce(b, ys, full=full)
        assert abs(lr.item() - lo.item()) < 1e-6
        assert torch.allclose(torch.autograd.grad(lr, a)[0], torch.autograd.grad(lo, b)[0], atol=1e-7)
        a1 = torch.randn(3, 1, 12, 20, generator=g).requires_grad_(); b1 = a1.detach().clone().requires_grad_()
        y1 = ys[:, 1:].contiguous()
        lr = ref.WeightedPartialCE(1, manual=True)(a1, y1.clone(), full=full); lo = O.weighted_partial_ce_bce(b1, y1, full=full)
        assert abs(lr.item() - lo.item()) < 1e-6
        assert torch.allclose(torch.autograd.grad(lr, a1)[0], torch.autograd.grad(lo, b1)[0], atol=1e-7)
    with pytest.raises(Exception):
        ref.WeightedPartialCE(3)(torch.rand(2, 3, 4, 4), torch.zeros(2, 3, 4, 4))      # only two classes are well-formed there


def test_known_answers_reference_and_oracle(ref):
    # closed-form known answers listed in SURVEY.md §8c
    ys = torch.zeros(1, 2, 2, 2); ys[0, 0, 0, 0] = 1; ys[0, 1, 0, 1] = 1; ys[0, 1, 1, 0] = 1
    yh = torch.full((1, 2, 2, 2), 0.5)
    for fn in (lambda a, b, **k: ref.WeightedPartialCE(2, manual=True)(a, b, **k),
               lambda a, b, **k: O.weighted_partial_ce(a, b, 2, **k)):
        assert abs(fn(yh, ys.clone()).item() - 1.0397208) < 1e-6
        assert abs(fn(yh, ys.clone(), reduction='sum').item() - 4.1588831) < 1e-5
    p = torch.full((2, 2, 2, 2), 0.5); t = torch.zeros(2, 2, 2, 2); t[:, 1] = 1
    assert abs(ref.DiceLoss()(p, t).item() - 0.5) < 1e-6 and abs(O.dice_loss(p, t).item() - 0.5) < 1e-6
    b = torch.zeros(1, 2, 4, 4); b[:, 0] = .8; b[:, 1] = .2
    q1 = torch.full((1, 2, 2, 2), .5); q2 = torch.zeros(1, 2, 1, 1); q2[:, 0] = .25; q2[:, 1] = .75
    assert abs(ref.InterlayerDivergence()([b, q1, q2]).item() - 0.4294572) < 1e-6
    assert abs(O.interlayer_divergence([b, q1, q2]).item() - 0.4294572) < 1e-6
    assert abs(ref.InterlayerDivergence()([b, q1, q2], weights=[2, 0]).item() + 0.5004024) < 1e-6
    assert abs(O.interlayer_divergence([b, q1, q2], weights=[2, 0]).item() + 0.5004024) < 1e-6
    assert abs(ref.InterlayerDivergence(divergence='JSD')([b, q1, q2]).item() - 0.0967728) < 1e-6
    assert abs(O.interlayer_divergence_jsd([b, q1, q2]).item() - 0.0967728) < 1e-6
    r = torch.tensor([[.5], [1.5]]); f = torch.tensor([[-.5], [0.]])
    assert abs(ref.LSDiscriminatorialLoss()(r, f).item() - 0.4375) < 1e-7
    assert abs(O.ls_discriminator_loss(r, f).item() - 0.4375) < 1e-7
    assert abs(ref.LSGeneratorLoss()(f).item() - 0.8125) < 1e-7 and abs(O.ls_generator_loss(f).item() - 0.8125) < 1e-7


@pytest.mark.parametrize("shape", [(2, 2, 32, 48), (3, 4, 17, 23)])
def test_losses_match_reference(ref, shape):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(3)
    yh = torch.softmax(torch.randn(B, C, H, W, generator=g), 1).requires_grad_()
    lab = torch.randint(0, C + 2, (B, H, W), generator=g)
    ys = torch.stack([(lab == c).float() for c in range(C)], 1)
    for kw in ({}, {"reduction": "sum"}, {"full": True}):
        a = ref.WeightedPartialCE(C, manual=True)(yh, ys.clone(), **kw)
        b = O.weighted_partial_ce(yh, ys.clone(), C, **kw)
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
        ga, = torch.autograd.grad(a, yh); gb, = torch.autograd.grad(b, yh)
        torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ref.DiceLoss()(yh, ys), O.dice_loss(yh, ys), rtol=1e-6, atol=1e-7)
    atts = [torch.softmax(torch.randn(B, C, max(H >> k, 1), max(W >> k, 1), generator=g), 1) for k in range(4)]
    for w in (None, [1, 2, 0], [0.5, 1, 1, 3]):
        torch.testing.assert_close(ref.InterlayerDivergence()(atts, weights=w), O.interlayer_divergence(atts, weights=w),
                                   rtol=1e-5, atol=1e-6)
    # Jensen-Shannon branch (losses.py:154-169), values and gradients, with and without stop_gradient
    for w in (None, [1, 2, 0]):
        for sg in (False, True):
            ar = [a.clone().requires_grad_() for a in atts]; ao = [a.clone().requires_grad_() for a in atts]
            lr = ref.InterlayerDivergence(divergence='JSD', stop_gradient=sg)(ar, weights=w)
            lo = O.interlayer_divergence_jsd(ao, weights=w, stop_gradient=sg)
            torch.testing.assert_close(lr, lo, rtol=1e-5, atol=1e-7)
            gr = torch.autograd.grad(lr, ar, allow_unused=True); go = torch.autograd.grad(lo, ao, allow_unused=True)
            for x, y in zip(gr, go):
                assert (x is None) == (y is None)
                if x is not None:
                    torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-8)


def test_segmentor_and_discriminator_match_reference(ref):
    torch.manual_seed(0)
    H = W = 80  # H/16 = 5 is odd: exercises the pad/crop path (compose.py:125-147)
    net = ref.OctaScribbleNet(torch.Size((2, 3, H, W)), torch.Size((2, 2, H, W)), True, False,
                              instance_noise=False, label_noise=False)
    net.train()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x, ys, _ = synth.octa_batch(2, H, W, seed=5)
    seg_sd = {k[len("segmentor."):]: v for k, v in sd.items() if k.startswith("segmentor.")}
    st = O.BNState()
    att_o, agg_o, x4_o = O.segmentor_forward(seg_sd, x, training=True, st=st)
    att_r, agg_r, x4_r = net.segmentor(x)
    torch.testing.assert_close(agg_o, agg_r, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x4_o, x4_r, rtol=1e-4, atol=1e-5)
    for a, b in zip(att_o, att_r):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    new_sd = net.segmentor.state_dict()
    for k, v in st.updated.items():
        torch.testing.assert_close(v, new_sd[k], rtol=1e-4, atol=1e-6, msg=k)
    # discriminator: one training forward advances u/v once
    dis_sd = {k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")}
    upd = {}
    pyr = synth.mask_pyramid(2, H, W)
    lo = O.discriminator_forward(dis_sd, pyr, depth=4, training=True, updated=upd)
    lr = net.discriminator(pyr)
    torch.testing.assert_close(lo, lr, rtol=1e-4, atol=1e-6)
    new_d = net.discriminator.state_dict()
    for k, v in upd.items():
        torch.testing.assert_close(v, new_d[k], rtol=1e-5, atol=1e-7, msg=k)
    # eval mode
    net.eval()
    att_r, agg_r, _ = net.segmentor(x)
    att_o, agg_o, _ = O.segmentor_forward({k: v for k, v in net.segmentor.state_dict().items()}, x, training=False)
    torch.testing.assert_close(agg_o, agg_r, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("gating_level", [None, 3, 4])
def test_parallel_head_oracle_matches_reference(ref, gating_level):
    """O.parallel_head_forward vs ResnestUnetParallelHead / ...AttentionGate (compose.py:233-527), train mode incl. the
    BatchNorm running statistics, then eval mode.  80x64: H/16 = 5 is odd (pad/crop path)."""
    torch.manual_seed(1)
    if gating_level is None:
        net = ref.compose.ResnestUnetParallelHead(2, False)
    else:
        net = ref.compose.ResnestUnetParallelHeadAttentionGate(2, False, None, gating_level)
    x, _, _ = synth.octa_batch(2, 80, 64, seed=6)
    for training in (True, False):
        net.train(training)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        st = O.BNState()
        out_o = O.parallel_head_forward(sd, x, training, gating_level, st)
        with torch.no_grad():
            out_r = net(x)
        if gating_level is None:
            assert out_r.shape == (2, 2, 2, 80, 64)
            torch.testing.assert_close(out_o, out_r, rtol=1e-4, atol=1e-5)
        else:
            (a_o, c_o), g_o = out_o
            (a_r, c_r), g_r = out_r
            assert len(a_o) == len(a_r) == (5 if gating_level > 3 else 4) and len(c_o) == len(c_r) == 2
            torch.testing.assert_close(g_o, g_r, rtol=1e-4, atol=1e-5)
            for u, v in zip(a_o + c_o, a_r + c_r):
                torch.testing.assert_close(u, v, rtol=1e-4, atol=1e-6)
        if training:
            new_sd = net.state_dict()
            for k, v in st.updated.items():
                torch.testing.assert_close(v, new_sd[k], rtol=1e-4, atol=1e-6, msg=k)
