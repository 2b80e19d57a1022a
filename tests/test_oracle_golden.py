"""Oracle (CPU restatement) against the golden vectors produced by the real reference (oracle/make_golden.py)."""
import os

import pytest
import torch

from oracle import octave_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cases():
    return torch.load(os.path.join(GOLD, "losses.pt"))


@pytest.mark.parametrize("name", ["pyr32", "pyr48x80", "generic"])
def test_oracle_losses_match_golden(cases, name):
    c = cases[name]
    C = c["yhat"].shape[1]
    yh = c["yhat"].clone().requires_grad_()
    l = O.weighted_partial_ce(yh, c["ys"].clone(), C)
    torch.testing.assert_close(l, c["wpce"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(torch.autograd.grad(l, yh)[0], c["wpce_g"], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(O.weighted_partial_ce(yh, c["ys"].clone(), C, reduction="sum"), c["wpce_sum"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(O.weighted_partial_ce(yh, c["ys"].clone(), C, full=True), c["wpce_full"], rtol=1e-5, atol=1e-6)
    ysb = c["ys"].clone(); ysb[:, 0] = 0
    torch.testing.assert_close(O.weighted_partial_ce(yh, ysb, C), c["wpce_ignore_bg"], rtol=1e-5, atol=1e-6)
    assert torch.equal(ysb, c["ys_after_ignore_bg"])
    lg = c["logits"].clone().requires_grad_()
    l = O.weighted_partial_ce(torch.softmax(lg, 1), c["ys"].clone(), C)
    torch.testing.assert_close(torch.autograd.grad(l, lg)[0], c["wpce_logits_g"], rtol=1e-4, atol=1e-8)
    l = O.dice_loss(yh, c["full"])
    torch.testing.assert_close(l, c["dice"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(torch.autograd.grad(l, yh)[0], c["dice_g"], rtol=1e-5, atol=1e-9)
    att = [a.clone().requires_grad_() for a in c["att"]]
    l = O.interlayer_divergence(att)
    torch.testing.assert_close(l, c["kld"], rtol=1e-5, atol=1e-6)
    for g, gg in zip(torch.autograd.grad(l, att), c["kld_g"]):
        torch.testing.assert_close(g, gg, rtol=1e-5, atol=1e-8)
    l = O.interlayer_divergence(att, weights=c["kld_weights"])
    torch.testing.assert_close(l, c["kld_w"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.ls_generator_loss(c["d_fake"]), c["lsg"])
    torch.testing.assert_close(O.ls_discriminator_loss(c["d_real"], c["d_fake"]), c["lsd"])


def test_oracle_jsd_known_answer():
    """SURVEY.md 8c closed-form answer confirmed against the imported reference (JSD branch, losses.py:154-169)."""
    b = torch.zeros(1, 2, 4, 4); b[:, 0] = .8; b[:, 1] = .2
    q1 = torch.full((1, 2, 2, 2), .5); q2 = torch.zeros(1, 2, 1, 1); q2[:, 0] = .25; q2[:, 1] = .75
    assert abs(O.interlayer_divergence_jsd([b, q1, q2]).item() - 0.0967728) < 1e-6


# ---------------------------------------------------------------------------------------------------
# whole networks: outputs of the REAL reference (oracle/make_golden.py make_networks) for seeded weights.  Only the seed
# is stored: the host mirror's seeded construction equals the reference's (tests/test_host_mirror.py).
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def nets():
    return torch.load(os.path.join(GOLD, "networks.pt"))


def _close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_oracle_segmentor_and_discriminator_match_reference_golden(nets):
    from architectures.models.octa import OctaScribbleNet
    from tests import synth
    B, H, W = nets["B"], nets["H"], nets["W"]
    x, _, _ = synth.octa_batch(B, H, W, seed=nets["batch_seed"])
    torch.manual_seed(nets["seed"])
    m = OctaScribbleNet(torch.Size((B, 3, H, W)), torch.Size((B, 2, H, W)), True, False, instance_noise=False, label_noise=False)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    seg = {k[len("segmentor."):]: v for k, v in sd.items() if k.startswith("segmentor.")}
    st = O.BNState()
    att, agg, x4 = O.segmentor_forward(seg, x, training=True, st=st)
    g = nets["segmentor_train"]
    _close(agg, g["agg"]); _close(x4, g["x4"])
    assert torch.equal(agg.argmax(1), g["agg"].argmax(1))
    assert len(att) == len(g["att"]) == 5
    for a, b in zip(att, g["att"]):
        _close(a, b, atol=1e-6)
    for k, v in g["bn"].items():
        _close(st.updated[k], v, atol=1e-6)
    dis = {k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")}
    upd = {}
    logit = O.discriminator_forward(dis, synth.mask_pyramid(B, H, W), depth=4, training=True, updated=upd)
    _close(logit, nets["discriminator_train"]["logit"], atol=1e-6)
    for k, v in nets["discriminator_train"]["u"].items():
        _close(upd[k], v, rtol=1e-5, atol=1e-7)
    seg_eval = dict(seg); seg_eval.update(st.updated)
    att, agg, x4 = O.segmentor_forward(seg_eval, x, training=False)
    g = nets["segmentor_eval"]
    _close(agg, g["agg"]); _close(x4, g["x4"])
    for a, b in zip(att, g["att"]):
        _close(a, b, atol=1e-6)


@pytest.mark.parametrize("name,gating_level", [("parallel_head", None), ("parallel_head_ag3", 3)])
def test_oracle_parallel_heads_match_reference_golden(nets, name, gating_level):
    from architectures.segmentor import compose
    from tests import synth
    x, _, _ = synth.octa_batch(nets["B"], nets["H"], nets["W"], seed=nets["batch_seed"])
    torch.manual_seed(nets["seed"])
    m = compose.ResnestUnetParallelHead(2, False) if gating_level is None else \
        compose.ResnestUnetParallelHeadAttentionGate(2, False, None, gating_level)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out = O.parallel_head_forward(sd, x, True, gating_level, O.BNState())
    g = nets[name]
    if gating_level is None:
        _close(out, g["agg"])
    else:
        (a, c), agg = out
        _close(agg, g["agg"])
        for u, v in zip(list(a) + list(c), g["att"] + g["att_c"]):
            _close(u, v, atol=1e-6)
