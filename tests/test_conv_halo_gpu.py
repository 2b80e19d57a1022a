"""Narrow-layer 3x3 halo kernel (csrc/conv_halo.cu): bit-exact against the generic tcgen05 implicit-GEMM kernel on the
same packed operands (forward, fused BN statistics, data gradient, accumulate mode, groups, ragged extents), and within
weight gradient, bf16 tolerance of a CPU fp32 convolution of the same bf16-rounded inputs (the arithmetic of nn.Conv2d in
/root/reference/architectures/extra/resnest.py:22-29,326-334)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [  # B, H, W, cin, cout, groups
    (2, 64, 64, 64, 32, 1),
    (2, 40, 52, 32, 64, 1),
    (3, 33, 47, 64, 128, 2),      # ragged extents, two groups
    (2, 100, 100, 128, 64, 2),
    (1, 16, 16, 32, 32, 1),       # image smaller than one patch
    (2, 50, 50, 64, 64, 1),
]


def _run(shape, enabled):
    from octave_b200 import ops
    from octave_b200.ops import Act, ConvSpec, lib
    B, H, W, cin, cout, groups = shape
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    x = Act(torch.randn(B, H, W, cin, device=dev, generator=g).bfloat16(), B, H, W, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin // groups, 3, 3, device=dev, generator=g) * 0.05)
    spec = ConvSpec(w, None, cin, cout, 3, 1, 1, groups)
    lib.octave_conv_halo_config(int(enabled), 0)
    try:
        y, st = ops.conv_fwd(x, spec, want_stats=True)
        dx = ops.conv_dgrad(y, spec, H, W)
        dx2 = ops.conv_dgrad(y, spec, H, W, out=Act(dx.buf.clone(), B, H, W, cin), accumulate=True)
        dw, _ = ops.conv_wgrad(x, y, spec)
        torch.cuda.synchronize()
    finally:
        lib.octave_conv_halo_config(1, 0)
    return x, w, y.buf.float(), st.clone(), dx.buf.float(), dx2.buf.float(), dw.float()


@pytest.mark.parametrize("shape", SHAPES)
def test_halo_matches_generic_kernel_bit_exact(shape):
    from octave_b200 import config
    config.set_compute_dtype("bf16")
    _, _, y0, s0, d0, a0, w0 = _run(shape, False)
    _, _, y1, s1, d1, a1, w1 = _run(shape, True)
    assert torch.equal(y0, y1)
    assert torch.equal(d0, d1)
    assert torch.equal(a0, a1)
    assert torch.allclose(s0, s1, rtol=1e-6, atol=1e-6)     # fp32 partial sums are folded in a different order
    # weight gradient: the halo kernel pairs filter taps inside one MMA; fp32 accumulation order differs
    assert (w0 - w1).abs().max() <= 1e-5 * w0.abs().max()


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_halo_vs_cpu_conv(shape):
    from octave_b200 import config
    config.set_compute_dtype("bf16")
    B, H, W, cin, cout, groups = shape
    x, w, y, st, dx, _, dw = _run(shape, True)
    xr = x.buf.float().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_()
    wr = w.detach().bfloat16().float().cpu()
    ref = F.conv2d(xr, wr, None, 1, 1, 1, groups)
    got = y.cpu().permute(0, 3, 1, 2)
    assert (got - ref).abs().max() <= 2e-2 * ref.abs().max()          # bf16 output rounding
    # fused statistics = per-channel sum / sum of squares of the stored (bf16) outputs
    assert torch.allclose(st[:cout].cpu(), got.double().sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[cout:].cpu(), (got.double() ** 2).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    # data gradient for upstream gradient = the bf16 output itself
    wq = wr.clone().requires_grad_()
    ref2 = F.conv2d(xr, wq, None, 1, 1, 1, groups)
    gx, gw = torch.autograd.grad(ref2, (xr, wq), got.contiguous())
    gd = dx.cpu().permute(0, 3, 1, 2)
    assert (gd - gx).abs().max() <= 2e-2 * gx.abs().max()
    # weight gradient (paired-tap halo kernel): same bf16 operands, fp32 accumulation
    assert (dw.cpu() - gw).abs().max() <= 1e-3 * gw.abs().max()
