"""On-GPU input pipeline (octave_b200.data): shapes, value ranges, label sparsity, determinism, geometric consistency."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_synthetic_batch_properties_and_determinism():
    from octave_b200 import data
    a = data.OnDeviceOcta(4, 96, 128, "cuda", seed=3, rank=0, augment=0)
    x, ys, real = a.next()
    x0, ys0, real0 = x.clone(), ys.clone(), [r.clone() for r in real]
    assert x.shape == (4, 3, 96, 128) and ys.shape == (4, 2, 96, 128)
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0 and float(x.std()) > 0.05
    assert torch.equal(x[:, 0], x[:, 1]) and torch.equal(x[:, 0], x[:, 2])            # one plane replicated (stem takes 3 channels)
    assert set(torch.unique(ys).tolist()) <= {0.0, 1.0}
    assert float((ys.sum(1) > 1).sum()) == 0.0                                          # one-hot or all-zero (unlabelled)
    frac = ys.mean(dim=(0, 2, 3))
    assert 0.002 < float(frac[0]) < 0.08 and 0.002 < float(frac[1]) < 0.15, frac        # sparse scribbles
    # foreground scribbles lie on vessels, background scribbles off them
    assert float((ys[:, 1] * (1 - a.vessel.float())).sum()) == 0.0 and float((ys[:, 0] * a.vessel.float()).sum()) == 0.0
    for k, r in enumerate(real):
        assert r.shape == (4, 2, 96 >> k, 128 >> k) and torch.equal(r.sum(1), torch.ones_like(r[:, 0]))
        if k:
            assert torch.equal(r, real[0][:, :, ::2 ** k, ::2 ** k])                    # nearest (strided) downsampling of level 0
    b = data.OnDeviceOcta(4, 96, 128, "cuda", seed=3, rank=0, augment=0)
    x1, ys1, real1 = b.next()
    assert torch.equal(x1, x0) and torch.equal(ys1, ys0) and all(torch.equal(p, q) for p, q in zip(real1, real0))
    x2, _, _ = b.next()
    assert not torch.equal(x2, x0)                                                      # the next batch differs
    c = data.OnDeviceOcta(4, 96, 128, "cuda", seed=3, rank=1, augment=0)
    assert not torch.equal(c.next()[0], x0)                                             # another rank draws other data


def test_augmentation_moves_image_and_labels_together():
    from octave_b200 import data
    g = torch.Generator().manual_seed(0)
    x = torch.rand(8, 3, 64, 64, generator=g).cuda()
    ys = (torch.rand(8, 2, 64, 64, generator=g) < 0.1).float().cuda()
    geo = data.AUG_FLIP_H | data.AUG_FLIP_V | data.AUG_ROT90
    xo, yo = data.augment(x, ys, seed=11, flags=geo)
    seen = set()
    for b in range(8):
        hit = None
        for rot in range(4):
            for fx in (0, 1):
                for fy in (0, 1):
                    # destination <- source mapping of the kernel: rotate, then flips, expressed on the source tensors
                    cand_x, cand_y = x[b], ys[b]
                    if fy: cand_x, cand_y = cand_x.flip(1), cand_y.flip(1)
                    if fx: cand_x, cand_y = cand_x.flip(2), cand_y.flip(2)
                    cand_x, cand_y = torch.rot90(cand_x, rot, (1, 2)), torch.rot90(cand_y, rot, (1, 2))
                    if torch.equal(cand_x, xo[b]) and torch.equal(cand_y, yo[b]):
                        hit = (rot, fx, fy)
        assert hit is not None, f"sample {b}: output is not a flip/rotation of the input applied to image and labels alike"
        seen.add(hit)
    assert len(seen) > 1                                                                # per-sample draws differ
    xp, yp = data.augment(x, ys, seed=11, flags=data.AUG_PHOTO)
    assert torch.equal(yp, ys) and not torch.equal(xp, x) and float(xp.min()) >= 0 and float(xp.max()) <= 1
    assert float((xp - x).abs().mean()) < 0.2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        data.augment(x.cpu(), ys.cpu(), seed=0)
