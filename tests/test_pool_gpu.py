"""K6 pools and the ConvTranspose gradient re-layout (csrc/pool_head.cu) against torch's CPU MaxPool2d / AvgPool2d autograd
and a plain index restatement — the pool configurations of /root/reference/architectures/extra/resnest.py:189 (avd
AvgPool2d(3, 2, padding=1)), :340 (MaxPool2d(3, 2, 1)), :383 (AvgPool2d(2, 2, ceil_mode=True, count_include_pad=False)),
plus a stride-1 window that takes the generic (non-batched) backward."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

POOLS = [  # kind, k, stride, pad, ceil_mode, count_include_pad
    ("max", 3, 2, 1, False, True),
    ("avg", 3, 2, 1, False, True),
    ("avg", 2, 2, 0, True, False),
    ("max", 3, 1, 1, False, True),
    ("avg", 3, 1, 1, False, False),
]
SIZES = [(2, 16, 20, 24), (3, 64, 25, 25), (1, 8, 7, 9), (2, 32, 50, 38)]   # B, C, H, W


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pool", POOLS)
def test_pool_fwd_bwd_vs_torch(pool, dtype):
    from octave_b200 import ops
    from octave_b200.ops import Act
    kind, k, s, p, ceil, cip = pool
    pd = ops.pool_desc(kind, k, s, p, ceil, cip)
    dev = torch.device("cuda")
    for (B, Cc, H, W) in SIZES:
        g = torch.Generator().manual_seed(B * 1000 + H)
        x = torch.randn(B, Cc, H, W, generator=g).to(dtype).float()          # values exactly representable in dtype
        xa = Act(x.permute(0, 2, 3, 1).contiguous().to(dtype).to(dev), B, H, W, Cc)
        y, arg = ops.pool_fwd(pd, xa)
        xr = x.clone().requires_grad_()
        if kind == "max":
            ref = F.max_pool2d(xr, k, s, p, ceil_mode=ceil)
        else:
            ref = F.avg_pool2d(xr, k, s, p, ceil_mode=ceil, count_include_pad=cip)
        assert tuple(ref.shape[2:]) == (y.H, y.W)
        got = y.buf.float().cpu().permute(0, 3, 1, 2)
        tol = 0.0 if (kind == "max" or dtype == torch.float32) else 8e-3
        assert (got - ref.detach()).abs().max() <= tol * max(1.0, float(ref.abs().max())) + (1e-6 if kind == "avg" else 0.0)
        dyv = torch.randn(ref.shape, generator=g).to(dtype).float()
        (gx,) = torch.autograd.grad(ref, (xr,), dyv)
        dya = Act(dyv.permute(0, 2, 3, 1).contiguous().to(dtype).to(dev), B, y.H, y.W, Cc)
        dx = ops.pool_bwd(pd, dya, arg, H, W)
        gd = dx.buf.float().cpu().permute(0, 3, 1, 2)
        tolb = 1e-6 if dtype == torch.float32 else 8e-3
        assert (gd - gx).abs().max() <= tolb * max(1.0, float(gx.abs().max())), (pool, (B, Cc, H, W))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 32, 16, 20, 8, 10), (3, 64, 25, 25, 13, 13), (1, 512, 9, 7, 5, 4), (2, 256, 50, 50, 25, 25)])
def test_space_to_depth_with_channel_sum(shape, dtype):
    """dst[b, h, w, t*C + c] = src[b, 2h + t//2, 2w + t%2, c] (zero outside src), chan_sum[c] = sum of src[..., c]."""
    from octave_b200 import ops
    from octave_b200.ops import Act
    B, Cc, Hs, Ws, H, W = shape
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Hs * 31 + Cc)
    src = torch.randn(B, Hs, Ws, Cc, generator=g).to(dtype)
    dst, cs = ops.space_to_depth(Act(src.to(dev), B, Hs, Ws, Cc), H, W, want_chan_sum=True)
    pad = torch.zeros(B, 2 * H, 2 * W, Cc, dtype=dtype)
    pad[:, :Hs, :Ws] = src
    ref = pad.reshape(B, H, 2, W, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, 4 * Cc)
    assert torch.equal(dst.buf.cpu(), ref)
    want = src.double().sum(dim=(0, 1, 2))
    assert torch.allclose(cs.double().cpu()[:Cc], want, rtol=1e-5, atol=1e-3)
    plain = ops.space_to_depth(Act(src.to(dev), B, Hs, Ws, Cc), H, W)
    assert torch.equal(plain.buf.cpu(), ref)
