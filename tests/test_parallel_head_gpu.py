"""Sibling segmentors with a second decoder head (octave_b200/network_parallel.py) against the CPU oracle
(O.parallel_head_forward, pinned to /root/reference/architectures/segmentor/compose.py:233-527 by
tests/test_oracle_vs_reference.py): train-mode forward + BatchNorm buffers, eval-mode gradients of every parameter
(x_1 and x_0_0 collect three gradients each), bf16 forward, folded inference and predict()."""
import pytest
import torch

from oracle import octave_oracle as O
from tests import synth
from tests.test_segmentor_gpu import _randomize_bn, l2err

pytestmark = pytest.mark.gpu


def _build(mode, gating_level, seed=0):
    from octave_b200 import config, network_parallel as npar
    config.set_compute_dtype(mode)
    torch.manual_seed(seed)
    if gating_level is None:
        net = npar.ResnestUnetParallelHead(2, False)
    else:
        net = npar.ResnestUnetParallelHeadAttentionGate(2, False, None, gating_level)
    return net.cuda().train()


def _flat(out, gated):
    """-> (attentions + attentions_c, agg)"""
    if not gated:
        return [], out
    (a, c), agg = out
    return list(a) + list(c), agg


def _loss(atts, agg, ys):
    """a scalar that reaches both heads and every attention map"""
    p = torch.softmax(agg, 2)
    loss = O.weighted_partial_ce(p[0], ys, 2) + 0.7 * O.weighted_partial_ce(p[1], ys, 2)
    for i, a in enumerate(atts):
        loss = loss + 0.05 * (i + 1) * (a * a).mean()
    return loss


@pytest.mark.parametrize("gating_level", [None, 3, 4])
def test_parallel_head_fp32_train_forward_and_buffers(gating_level):
    net = _build("fp32", gating_level, seed=1)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x, _, _ = synth.octa_batch(4, 80, 64, seed=7)          # 80/16 = 5: pad/crop path
    st = O.BNState()
    atts_o, agg_o = _flat(O.parallel_head_forward(sd, x, True, gating_level, st), gating_level is not None)
    atts, agg = _flat(net(x.cuda()), gating_level is not None)
    assert agg.shape == agg_o.shape == (2, 4, 2, 80, 64)
    assert l2err(agg, agg_o) < 1e-4, l2err(agg, agg_o)
    assert torch.equal(agg.argmax(2).cpu(), agg_o.argmax(2))
    assert len(atts) == len(atts_o) == {None: 0, 3: 6, 4: 7}[gating_level]
    for a, b in zip(atts, atts_o):
        assert a.shape == b.shape and l2err(a, b) < 1e-4
    new = net.state_dict()
    for k, v in st.updated.items():
        if "num_batches" in k:
            assert int(new[k]) == int(v)
        else:
            assert l2err(new[k], v) < 1e-4, k


@pytest.mark.parametrize("gating_level", [None, 3])
def test_parallel_head_fp32_eval_gradients(gating_level):
    net = _build("fp32", gating_level, seed=2)
    _randomize_bn(net)
    net.eval()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x, ys, _ = synth.octa_batch(2, 64, 80, seed=8)
    gated = gating_level is not None
    sdr = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    atts_o, agg_o = _flat(O.parallel_head_forward(sdr, x, False, gating_level), gated)
    names = [k for k, v in sdr.items() if v.requires_grad]
    g_o = dict(zip(names, torch.autograd.grad(_loss(atts_o, agg_o, ys), [sdr[k] for k in names], allow_unused=True)))
    atts, agg = _flat(net(x.cuda()), gated)
    assert l2err(agg, agg_o) < 1e-4
    _loss(atts, agg, ys.cuda()).backward()
    params = dict(net.named_parameters())
    gscale = max(float(v.abs().max()) for v in g_o.values() if v is not None)
    errs = []
    for k, go in g_o.items():
        if go is None:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
            continue
        if float(go.abs().max()) < 1e-7 * gscale:
            continue
        assert params[k].grad is not None, k
        errs.append((l2err(params[k].grad, go), k))
    errs.sort(reverse=True)
    print("worst parallel-head gradient errors:", errs[:6], "median", errs[len(errs) // 2])
    assert errs[0][0] < 5e-3, errs[:5]
    assert errs[len(errs) // 2][0] < 5e-4
    # only the main head in the loss: the parallel branch gets no gradient, the shared encoder still does
    net.zero_grad(set_to_none=True)
    _, agg = _flat(net(x.cuda()), gated)
    O.weighted_partial_ce(torch.softmax(agg[0], 1), ys.cuda(), 2).backward()
    for k in ("fc_c.weight", "decoder_1_c.conv.0.weight", "upsampling_0_c.up.weight"):
        assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
    for k in ("fc.weight", "decoder_1.conv.0.weight", "encoder_1.0.conv1.weight"):
        assert params[k].grad is not None and float(params[k].grad.abs().max()) > 0.0, k


def test_parallel_head_bf16_forward_and_inference():
    """bf16 mode.  Train-mode step runs through both branches (finite outputs and gradients; train-mode parity at test
    sizes is ill-conditioned, see test_segmentor_gpu.py); eval-mode forward, folded inference and predict() against the
    fp32 oracle."""
    from octave_b200 import config
    net = _build("bf16", 3, seed=3)
    _randomize_bn(net, seed=9)
    x, ys, _ = synth.octa_batch(2, 96, 64, seed=10)
    atts, agg = _flat(net(x.cuda()), True)
    assert bool(torch.isfinite(agg).all())
    _loss(atts, agg, ys.cuda()).backward()
    params = dict(net.named_parameters())
    for k, p in params.items():
        if k.startswith("aag_4."):
            assert p.grad is None            # level 4 is not gated at gating_level 3 (compose.py:465)
        else:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    net.eval()
    _randomize_bn(net, seed=9)               # the train-mode step above moved the running statistics
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    atts_e, agg_e = _flat(O.parallel_head_forward(sd, x, False, 3), True)
    with torch.no_grad():
        (a, c), pred = net.predict(x.cuda(), method='one-hot')      # folded conv+BN inference path
        config.fold_bn_inference = False
        try:
            _, agg_u = net(x.cuda())
        finally:
            config.fold_bn_inference = True
        (a_f, c_f), agg_f = net(x.cuda())
    assert len(a) == 4 and len(c) == 2 and pred.shape[:2] == (2, 2) and pred.shape[-2:] == (96, 64)
    assert l2err(agg_f, agg_e) < 3e-2 and l2err(agg_f, agg_u) < 3e-2, (l2err(agg_f, agg_e), l2err(agg_f, agg_u))
    for u, v in zip(list(a_f) + list(c_f), atts_e):
        assert u.shape == v.shape and l2err(u, v) < 5e-2
