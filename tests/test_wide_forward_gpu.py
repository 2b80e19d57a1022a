"""Layers that take the 128 x 256 forward tile kernel (csrc/conv_tc.cu conv_tc_kernel<256,64,4>, slab epilogue): output,
fused BatchNorm statistics and data gradient against a CPU fp32 convolution of the same bf16 operands — the arithmetic of
nn.Conv2d in /root/reference/architectures/extra/resnest.py:22-29.  (The kernel's bit-exactness against its 3-stage
predecessor is recorded in profiles/slab_epilogue_ab_r01.log.)"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(24, 37, 41, 256, 512, 3), (80, 32, 32, 1024, 256, 1)])
def test_wide_forward_256_column_tiles_vs_cpu_conv(shape):
    """Forward / data gradient of layers that take the 128 x 256 tile kernel (conv_tc_kernel<256,64,4>: >= 4 waves of
    tiles, K >= 1024): output, fused BatchNorm statistics and dgrad against a CPU fp32 convolution of the same bf16
    operands.  Same bounds as tests/test_conv_halo_gpu.py::test_halo_vs_cpu_conv."""
    from octave_b200 import config, ops
    from octave_b200.ops import Act, ConvSpec
    config.set_compute_dtype("bf16")
    B, H, W, cin, cout, k = shape
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    x = Act(torch.randn(B, H, W, cin, device=dev, generator=g).bfloat16(), B, H, W, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.03)
    spec = ConvSpec(w, None, cin, cout, k, 1, k // 2, 1)
    y, st = ops.conv_fwd(x, spec, want_stats=True)
    dx = ops.conv_dgrad(y, spec, H, W)
    torch.cuda.synchronize()
    xr = x.buf.float().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_()
    wr = w.detach().bfloat16().float().cpu()
    ref = F.conv2d(xr, wr, None, 1, k // 2)
    got = y.buf.float().cpu().permute(0, 3, 1, 2)
    assert (got - ref.detach()).abs().max() <= 2e-2 * ref.abs().max()
    assert torch.allclose(st[:cout].cpu(), got.double().sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(st[cout:].cpu(), (got.double() ** 2).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-2)
    (gx,) = torch.autograd.grad(ref, (xr,), got.contiguous())
    gd = dx.buf.float().cpu().permute(0, 3, 1, 2)
    assert (gd - gx).abs().max() <= 2e-2 * gx.abs().max()
