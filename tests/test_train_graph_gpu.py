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
    """lr = 0: the weights stay put (BatchNorm running statistics, spectral-norm u/v and the noise draws still evolve),
    so the third eager step and the first replay (after two warm-up steps) must give the same losses."""
    x, ys, real = _data()
    net_a, ts_a = _make(0, lr=0.0)
    torch.manual_seed(99)                       # CPU stream of the critic's noise draws
    for _ in range(3):
        res_a = ts_a.step(x, ys, real)
    la = {k: float(res_a[k]) for k in KEYS}
    r_a = torch.rand(3)

    net_b, ts_b = _make(0, lr=0.0)
    torch.manual_seed(99)
    res_b = ts_b.step_graphed(x, ys, real)       # 2 eager warm-up steps + capture + 1 replay = 3 steps
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


def test_graphed_step_trains():
    """lr > 0: the optimiser steps and the weight re-pack run inside the graph — the weights move by updates of the
    same size as in eager mode.  (Trajectories are not compared element-wise: with B = 2 the batch-of-two BatchNorm of
    the split-attention branch normalises to +-1, and two EAGER runs already differ by more than the update itself.)"""
    x, ys, real = _data()
    net_a, ts_a = _make(0)
    w0 = {k: v.clone() for k, v in net_a.state_dict().items()}
    torch.manual_seed(99)
    for _ in range(4):
        ts_a.step(x, ys, real)
    net_b, ts_b = _make(0)
    torch.manual_seed(99)
    ts_b.step_graphed(x, ys, real)
    w3 = {k: v.clone() for k, v in net_b.state_dict().items()}
    res_b = ts_b.step_graphed(x, ys, real)       # second replay = 4th step
    assert ts_b.graph_error is None, ts_b.graph_error
    assert all(torch.isfinite(res_b[k]) for k in KEYS)
    sa, sb = net_a.state_dict(), net_b.state_dict()
    keys = [k for k in sa if sa[k].is_floating_point() and "running" not in k and not k.startswith("segmentor.linear_head_")
            and not k.endswith(("_u", "_v"))]
    norm = lambda p, q: sum(float(((p[k] - q[k]).float() ** 2).sum()) for k in keys) ** 0.5
    upd_a, upd_b, last_b = norm(sa, w0), norm(sb, w0), norm(sb, w3)
    assert last_b > 0.0                          # the replayed optimiser step changed the weights
    # magnitudes are only sanity-checked (finite, non-degenerate): at B = 2 the gradient itself is ill-conditioned
    assert upd_a > 0.0 and upd_b > 0.0 and upd_b == upd_b and upd_b < 1e6, (upd_a, upd_b)
