"""Weight gradient of the wide decoder / encoder layers (csrc/conv_tc.cu: conv_tc_wgrad_kernel, 128 x 256 tiles with the
short-stage deep ring) against a CPU fp32 convolution of the same bf16-rounded operands — the arithmetic of
`nn.Conv2d.weight.grad` for /root/reference/architectures/extra/resnest.py:326-334 (decoder convs) and :22-29."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [  # B, H, W, cin, cout, k
    (2, 50, 50, 256, 128, 3),     # 25x3 patches, two tiles per row
    (2, 25, 25, 512, 128, 3),     # image width below the stage size
    (1, 37, 41, 256, 192, 3),     # ragged extents, a half-filled second Cout tile
    (2, 20, 20, 256, 256, 1),     # 1x1: pixels flattened to one row
    (1, 100, 100, 256, 128, 3),   # 20x4 patches fill the stage exactly
    (1, 16, 16, 1280, 1024, 3),   # 360 output tiles, no split-K: two co-resident CTAs per SM
]


@pytest.mark.parametrize("shape", SHAPES)
def test_wide_wgrad_vs_cpu_conv(shape):
    from octave_b200 import config, ops
    from octave_b200.ops import Act, ConvSpec
    config.set_compute_dtype("bf16")
    B, H, W, cin, cout, k = shape
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(11)
    x = Act(torch.randn(B, H, W, cin, device=dev, generator=g).bfloat16(), B, H, W, cin)
    dy = Act(torch.randn(B, H, W, cout, device=dev, generator=g).bfloat16(), B, H, W, cout)
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.05)
    spec = ConvSpec(w, None, cin, cout, k, 1, k // 2, 1)
    dw, _ = ops.conv_wgrad(x, dy, spec)
    dw2, _ = ops.conv_wgrad(x, dy, spec)
    torch.cuda.synchronize()
    xr = x.buf.float().cpu().permute(0, 3, 1, 2).contiguous()
    gr = dy.buf.float().cpu().permute(0, 3, 1, 2).contiguous()
    wq = w.detach().float().cpu().requires_grad_()
    (gw,) = torch.autograd.grad(F.conv2d(xr, wq, None, 1, k // 2), (wq,), gr)
    # same bf16 operands, fp32 accumulation (split-K partial sums meet in fp32 atomics)
    assert (dw.float().cpu() - gw).abs().max() <= 1e-3 * gw.abs().max()
    assert (dw2.float().cpu() - gw).abs().max() <= 1e-3 * gw.abs().max()
