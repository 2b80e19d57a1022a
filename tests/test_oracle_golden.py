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
