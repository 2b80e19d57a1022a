"""The split-attention branch of SplAtConv2d (fc1 -> bn1 -> relu -> fc2 -> r-softmax, reference extra/resnest.py:116-127) as
two launches per direction (octave_glinear_bn_relu_fwd, octave_glinear_rsoftmax_fwd, octave_rsoftmax_glinear_bn_bwd) against
the four separate kernels: same summation orders, so the results must be BIT-identical (the separate kernels are the ones the
oracle parity tests of the blocks cover, and what batches larger than 32 still use)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,inter,training", [(8, 64, 32, True), (32, 256, 128, True), (5, 128, 64, True), (32, 512, 256, False)])
def test_fused_attention_branch_is_bit_identical_to_the_separate_kernels(B, C, inter, training):
    from octave_b200 import config, ops
    assert ops.attn_fused_ok(B, C, inter, 1, 2)
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    gap, w1, b1 = r(B, C) * 50.0, r(inter, C) * 0.1, r(inter) * 0.1
    w2, b2 = r(2 * C, inter) * 0.1, r(2 * C) * 0.1
    gamma, beta = torch.rand(inter, device="cuda", generator=g) + 0.5, r(inter) * 0.1
    scale = 1.0 / 37.0

    def stats():
        return (torch.full((inter,), 0.25, device="cuda"), torch.full((inter,), 1.5, device="cuda"),
                torch.zeros((), dtype=torch.int64, device="cuda"))

    # forward ---------------------------------------------------------------------------------------
    rm_a, rv_a, nbt_a = stats()
    h1_a = ops.glinear_fwd(gap, w1, b1, 1, scale)
    h1n_a, mi_a = ops.bn1d_relu_fwd(h1_a, gamma, beta, rm_a, rv_a, nbt_a if training else None, 1e-5, 0.1, training)
    att_a = ops.rsoftmax_fwd(ops.glinear_fwd(h1n_a, w2, b2, 1, 1.0), 2)
    rm_b, rv_b, nbt_b = stats()
    h1_b, h1n_b, mi_b = ops.glinear_bn_relu_fwd(gap, w1, b1, scale, gamma, beta, rm_b, rv_b, nbt_b if training else None, 1e-5, 0.1, training)
    att_b = ops.glinear_rsoftmax_fwd(h1n_b, w2, b2, C)
    for name, a, b in (("h1", h1_a, h1_b), ("h1n", h1n_a, h1n_b), ("mean_invstd", mi_a, mi_b), ("att", att_a, att_b),
                       ("running_mean", rm_a, rm_b), ("running_var", rv_a, rv_b), ("num_batches_tracked", nbt_a, nbt_b)):
        assert torch.equal(a, b), name
    assert int(nbt_b) == (1 if training else 0)
    assert (h1n_b > 0).any() and (h1n_b == 0).any()           # the ReLU gate is exercised both ways
    torch.testing.assert_close(att_b[:, :C] + att_b[:, C:], torch.ones(B, C, device="cuda"), rtol=1e-6, atol=1e-6)

    # backward (one writer per output in the separate data-gradient kernel, as in the fused one) -----------------
    datt = r(B, 2 * C)
    config.set_deterministic(True)
    try:
        dlogits = ops.rsoftmax_bwd(datt, att_a, 2)
        dh1n, dw2_a, db2_a = ops.glinear_bwd(dlogits, h1n_a, w2, 1, 1.0)
        dx_a, dg_a, dbt_a = ops.bn1d_relu_bwd(dh1n, h1_a, h1n_a, gamma, mi_a, training)
    finally:
        config.set_deterministic(False)
    dx_b, dg_b, dbt_b, dw2_b, db2_b = ops.attn_bwd_fused(datt, att_b, w2, h1_b, h1n_b, gamma, mi_b, training, 1)
    for name, a, b in (("dx", dx_a, dx_b), ("dgamma", dg_a, dg_b), ("dbeta", dbt_a, dbt_b), ("dw2", dw2_a, dw2_b), ("db2", db2_a, db2_b)):
        assert torch.equal(a, b), name
    assert torch.isfinite(dx_b).all() and float(dx_b.abs().max()) > 0


def test_large_batches_keep_the_separate_kernels():
    from octave_b200 import ops
    assert not ops.attn_fused_ok(64, 256, 128, 1, 2)           # two 32-row slabs: BatchNorm1d needs the whole batch in one warp
    assert not ops.attn_fused_ok(32, 256, 128, 2, 2)           # grouped linears: other column layout
