"""DiscriminatorBlock on the kernels vs the CPU oracle: logits, gradients w.r.t. the input maps (G-step) and the
parameters incl. spectral-norm `weight_orig` (D-step), power-iteration buffers, instance/label noise RNG parity."""
import pytest
import torch

from oracle import octave_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu


def l2err(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("mode,tol,gtol", [("fp32", 1e-4, 1e-3), ("bf16", 1e-2, 8e-2)])
@pytest.mark.parametrize("noise", [False, True])
def test_discriminator_parity(mode, tol, gtol, noise):
    from octave_b200 import config, discriminator, losses
    config.set_compute_dtype(mode)
    B, H = 3, 96
    torch.manual_seed(0)
    D = discriminator.DiscriminatorBlock(torch.Size((B, 2, H, H)), True, depth=4, num_filters=64,
                                         instance_noise=noise, label_noise=noise)
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    D = D.cuda().train()
    maps = [m.clone() for m in synth.prob_maps(B, 2, H, H, 5, seed=2)]
    # oracle: replay the same CPU RNG draws
    torch.manual_seed(123)
    noise_plane = torch.normal(mean=.0, std=.2, size=(H, H)) if noise else None
    flip = bool(torch.FloatTensor(1).uniform_(0, 1) < 0.1) if noise else False
    sdr = {k: (v.clone().requires_grad_() if v.is_floating_point() and not k.endswith(("_u", "_v")) else v.clone()) for k, v in sd.items()}
    mo = [m.clone().requires_grad_() for m in maps]
    upd = {}
    lo = O.discriminator_forward(sdr, mo, depth=4, training=True, is_training_flag=True, noise=noise_plane,
                                 instance_noise=noise, flip=flip, updated=upd)
    loss_o = O.ls_generator_loss(lo)
    names = [k for k, v in sdr.items() if v.requires_grad]
    go = torch.autograd.grad(loss_o, mo + [sdr[k] for k in names])
    torch.manual_seed(123)
    mc = [m.cuda().requires_grad_() for m in maps]
    lc = D(mc)
    assert lc.shape == (B, 1)
    assert l2err(lc, lo) < tol, l2err(lc, lo)
    loss = losses.LSGeneratorLoss()(lc)
    assert abs(loss.item() - loss_o.item()) <= tol * abs(loss_o.item()) + 1e-6
    params = dict(D.named_parameters())
    gc = torch.autograd.grad(loss, mc + [params[k] for k in names])
    for nm, a, b in zip([f"map{k}" for k in range(5)] + names, gc, go):
        assert l2err(a, b) < gtol, (nm, l2err(a, b))
    new = D.state_dict()
    for k, v in upd.items():
        assert l2err(new[k], v) < 1e-4, k
    # RNG stream advanced identically (one normal plane + one uniform per call)
    if noise:
        torch.manual_seed(123)
        torch.normal(mean=.0, std=.2, size=(H, H)); torch.FloatTensor(1).uniform_(0, 1)
        expect = torch.rand(2)
        torch.manual_seed(123)
        D([m.cuda() for m in maps])
        assert torch.equal(torch.rand(2), expect)


def test_discriminator_d_step_and_frozen_g_step():
    from octave_b200 import config, discriminator, losses
    config.set_compute_dtype("fp32")
    B, H = 2, 64
    torch.manual_seed(1)
    D = discriminator.DiscriminatorBlock(torch.Size((B, 2, H, H)), True, depth=4, instance_noise=False, label_noise=False).cuda().train()
    real = [m.cuda() for m in synth.mask_pyramid(B, H, H)]
    fake = [m.cuda().requires_grad_() for m in synth.prob_maps(B, 2, H, H, 5, seed=4)]
    l = losses.LSDiscriminatorialLoss()(D(real), D([f.detach() for f in fake]))
    l.backward()
    assert all(p.grad is not None for p in D.parameters())
    for p in D.parameters():
        p.requires_grad_(False)
    g = losses.LSGeneratorLoss()(D(fake))
    g.backward()
    assert all(f.grad is not None and float(f.grad.abs().max()) > 0 for f in fake)
    with pytest.raises(Exception, match="depth"):
        D(real[:3])
