"""TrainStep.step_graphed (whole adversarial step replayed from one CUDA graph) against the eager TrainStep.step:
same seeds, same inputs, same CPU random stream for the critic's instance / label noise -> same losses and weights."""
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu


def _make(seed, lr=1e-3):
    from octave_b200 import config
    from octave_b200.model import OctaScribbleNet
    from octave_b200.train import TrainStep
    config.set_compute_dtype("bf16")
    config.nan_check = False
    B, H = 2, 64
    torch.manual_seed(seed)
    net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
    return net, TrainStep(net, lr=lr)


def _data():
    from octave_b200 import synth as osynth
    B, H = 2, 64
    x, ys, _ = osynth.octa_batch(B, H, H, seed=3, n_ridges=6)
    real = osynth.mask_pyramid(B, H, H, seed=5, n_ridges=6)
    return x.cuda(), ys.cuda(), [r.cuda() for r in real]


KEYS = ("supervised", "divergence", "generator", "discriminator")


def test_graphed_step_matches_eager_frozen_weights():
    """lr = 0.  Capture warms up with two eager steps and then restores every parameter / buffer, so the FIRST replay is one
    step from the initial state, fed with the third set of CPU noise draws.  The eager twin does the same by hand: two
    steps, buffers restored, third step -> same losses, same updated buffers, same CPU generator state."""
    x, ys, real = _data()
    net_a, ts_a = _make(0, lr=0.0)
    sd0 = {k: v.clone() for k, v in net_a.state_dict().items()}
    torch.manual_seed(99)                       # CPU stream of the critic's noise draws
    for _ in range(2):
        ts_a.step(x, ys, real)
    net_a.load_state_dict(sd0)
    res_a = ts_a.step(x, ys, real)
    la = {k: float(res_a[k]) for k in KEYS}
    r_a = torch.rand(3)

    net_b, ts_b = _make(0, lr=0.0)
    torch.manual_seed(99)
    res_b = ts_b.step_graphed(x, ys, real)       # 2 eager warm-up steps (state restored) + capture + 1 replay
    assert ts_b.graph_error is None, ts_b.graph_error
    assert ts_b._graph is not None and ts_b.graph_kernel_nodes > 500
    lb = {k: float(res_b[k]) for k in KEYS}
    # the CPU generator advanced identically (same number and order of instance / label noise draws)
    assert torch.equal(torch.rand(3), r_a)
    for k in KEYS:
        assert abs(la[k] - lb[k]) <= 2e-3 * abs(la[k]) + 1e-5, (k, la[k], lb[k])
    # buffers updated inside the graph: running statistics and power-iteration vectors match the eager run
    sa, sb = net_a.state_dict(), net_b.state_dict()
    for k in sa:
        if k.endswith(("running_mean", "running_var", "weight_u", "weight_v", "num_batches_tracked")):
            assert torch.allclose(sa[k].float(), sb[k].float(), rtol=2e-2, atol=1e-3), k
        if k.endswith("num_batches_tracked") and "linear_head_" not in k:
            assert int(sb[k]) == 1, (k, int(sb[k]))          # ONE step was applied, not three
    with pytest.raises(ValueError, match="static shapes"):
        ts_b.step_graphed(x[:1], ys[:1], [r[:1] for r in real])


def test_graphed_step_applies_the_sgd_update_exactly():
    """lr > 0: the optimiser steps and the weight re-pack run inside the graph.  The update rule is checked against the
    graph's own gradients (trajectories are not compared with an eager run element-wise: at B = 2 the batch-of-two
    BatchNorm of the split-attention branch makes the gradient chaotic, two eager runs already differ):
      first replay   w1 = w0 - lr * g1                (momentum buffer initialised with g1)
      second replay  w2 = w1 - lr * (0.9 * g1 + g2)"""
    x, ys, real = _data()
    lr = 1e-2
    net, ts = _make(0, lr=lr)
    params = {n: p for n, p in net.named_parameters() if not n.startswith("segmentor.linear_head_")}
    w0 = {n: p.detach().clone() for n, p in params.items()}
    torch.manual_seed(99)
    res = ts.step_graphed(x, ys, real)
    assert ts.graph_error is None, ts.graph_error
    torch.cuda.synchronize()
    assert all(torch.isfinite(res[k]) for k in KEYS)
    g1 = {n: p.grad.detach().clone() for n, p in params.items() if p.grad is not None}
    w1 = {n: p.detach().clone() for n, p in params.items()}
    assert len(g1) > 300
    moved = 0
    for n, g in g1.items():
        exp = w0[n] - lr * g
        assert torch.allclose(w1[n], exp, rtol=1e-5, atol=1e-8), n
        moved += int(not torch.equal(w1[n], w0[n]))
    assert moved > 300                            # the replayed optimiser step changed (nearly) every tensor
    ts.step_graphed(x, ys, real)
    torch.cuda.synchronize()
    for n, g in g1.items():
        g2 = params[n].grad
        exp = w1[n] - lr * (0.9 * g + g2)
        assert torch.allclose(params[n].detach(), exp, rtol=1e-4, atol=1e-7), n
