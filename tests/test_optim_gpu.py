"""Multi-tensor optimiser kernel (octave_optim_multi) against torch.optim.SGD / torch.optim.AdamW on the same parameters
and gradients: the update rules are torch's (the reference ships no optimiser, README.md:39-47)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]


SHAPES = [(64, 32, 3, 3), (7,), (256, 64, 1, 1), (13, 5), (1,), (2048, 1024, 3, 3), (33, 3, 3, 3)]


@pytest.mark.parametrize("kind", ["sgd", "sgd_wd", "adamw"])
def test_multi_tensor_step_matches_torch(kind):
    from octave_b200.optim import FusedAdamW, FusedSGD
    pa, pb = _params(0, SHAPES), _params(0, SHAPES)
    if kind == "adamw":
        oa, ob = torch.optim.AdamW(pa, lr=3e-3, weight_decay=0.05), FusedAdamW(pb, lr=3e-3, weight_decay=0.05)
    else:
        wd = 1e-2 if kind == "sgd_wd" else 0.0
        oa, ob = torch.optim.SGD(pa, lr=0.05, momentum=0.9, weight_decay=wd), FusedSGD(pb, lr=0.05, momentum=0.9, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda()
            if step == 2 and a.numel() == 7 and kind != "adamw":   # (FusedAdamW keeps ONE step count for its bias corrections)
                a.grad = b.grad = None                     # a parameter without a gradient is skipped, like torch
                continue
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
        for k, (a, b) in enumerate(zip(pa, pb)):
            torch.testing.assert_close(b.detach(), a.detach(), rtol=1e-5, atol=1e-6, msg=lambda m: f"{kind} step {step} tensor {k}: {m}")
    ob.zero_grad()
    assert all(p.grad is None for p in pb)


def test_train_step_uses_one_optimizer_launch_per_module():
    from octave_b200 import _lib, config
    from octave_b200.model import OctaScribbleNet
    from octave_b200.optim import FusedSGD
    from octave_b200.train import TrainStep
    from octave_b200 import synth
    config.set_compute_dtype("bf16"); config.nan_check = False
    torch.manual_seed(0)
    net = OctaScribbleNet(torch.Size((2, 3, 64, 64)), torch.Size((2, 2, 64, 64)), True, False, instance_noise=False, label_noise=False).cuda().train()
    ts = TrainStep(net, lr=1e-2)
    assert isinstance(ts.opt_g, FusedSGD) and isinstance(ts.opt_d, FusedSGD)
    x, ys, _ = synth.octa_batch(2, 64, 64, seed=0, n_ridges=6)
    real = [r.cuda() for r in synth.mask_pyramid(2, 64, 64, n_ridges=6)]
    w0 = net.segmentor.fc.weight.detach().clone()
    res = ts.step(x.cuda(), ys.cuda(), real)
    g = net.segmentor.fc.weight.grad
    torch.testing.assert_close(net.segmentor.fc.weight.detach(), w0 - 1e-2 * g, rtol=1e-6, atol=1e-8)
    assert torch.isfinite(res['total'])
    # the update invalidates the cached bf16 operand packs: the next forward must run on the NEW weights
    from octave_b200 import ops
    conv = net.segmentor.decoder_2.conv[0]
    v0 = conv.weight._version
    ts.step(x.cuda(), ys.cuda(), real)
    assert conv.weight._version > v0
    fresh = ops.ConvSpec(conv.weight.detach().clone(), None, conv.in_channels, conv.out_channels, 3, 1, 1, 1).pack(ops._lib_pack.FWD)
    assert torch.equal(conv._oct_spec.pack(ops._lib_pack.FWD), fresh)


def test_bf16_gradient_bucket_pack_and_unpack_round_trip():
    """octave_grad_pack_bf16 / octave_grad_unpack_bf16 (data-parallel bf16 buckets): the flat bucket holds exactly
    torch's round-to-nearest bf16 cast of every tensor at its 16-byte slot, the way back is the exact widening; sizes cover
    the scalar path (13, 1, odd), block boundaries (4096 k + 5) and more tensors than one launch carries (300 > 128)."""
    from octave_b200.train import GradAllReducer
    g = torch.Generator(device="cuda").manual_seed(3)
    sizes = [13, 1, 8, 4096 * 3 + 5, 8192, 77, 64 * 3 * 3 * 3] + [16 + (i % 7) for i in range(300)]
    grads = [torch.randn(n, device="cuda", generator=g) * (10.0 ** ((i % 5) - 2)) for i, n in enumerate(sizes)]
    keep = [t.clone() for t in grads]
    offs, off = [], 0
    for t in grads:
        offs.append(off)
        off += (t.numel() + 7) & ~7
    flat = torch.full((off,), float("nan"), dtype=torch.bfloat16, device="cuda")
    GradAllReducer._bucket_kernel("octave_grad_pack_bf16", grads, offs, flat)
    for t, o in zip(keep, offs):
        assert torch.equal(flat[o:o + t.numel()], t.bfloat16())
    for t in grads:
        t.fill_(-1.0)
    GradAllReducer._bucket_kernel("octave_grad_unpack_bf16", grads, offs, flat)
    for t, k in zip(grads, keep):
        assert torch.equal(t, k.bfloat16().float())
