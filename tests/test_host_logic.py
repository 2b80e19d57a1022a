"""Host-side logic that needs no GPU: conv+BatchNorm folding algebra and its cache (network._folded_spec, the inference
path of /root/reference/architectures/segmentor/compose.py:189-199 with every BatchNorm in eval mode), gate placement of
the parallel-head siblings (compose.py:465-507), dense-group merging of grouped convolutions (ops.ConvSpec)."""
import torch
import torch.nn.functional as F


def test_folded_conv_bn_equals_eval_batchnorm_of_conv():
    from octave_b200 import network
    torch.manual_seed(0)
    for bias in (False, True):
        conv = torch.nn.Conv2d(8, 16, 3, padding=1, bias=bias)
        bn = torch.nn.BatchNorm2d(16).eval()
        with torch.no_grad():
            bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5); bn.weight.normal_(); bn.bias.normal_()
        x = torch.randn(2, 8, 7, 9)
        spec = network._folded_spec(conv, bn)
        assert (spec.cin, spec.cout, spec.k, spec.stride, spec.pad, spec.groups) == (8, 16, 3, 1, 1, 1)
        with torch.no_grad():
            want = bn(conv(x))
            got = F.conv2d(x, spec.weight, spec.bias, 1, 1)
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
        assert network._folded_spec(conv, bn) is spec                      # cached
        with torch.no_grad():
            bn.running_var.mul_(2.0)                                       # a statistic moved: refold
        spec2 = network._folded_spec(conv, bn)
        assert spec2 is not spec
        with torch.no_grad():
            torch.testing.assert_close(F.conv2d(x, spec2.weight, spec2.bias, 1, 1), bn(conv(x)), rtol=1e-5, atol=1e-5)
        with torch.no_grad():
            conv.weight.add_(0.1)                                          # an optimiser step: refold
        assert network._folded_spec(conv, bn) is not spec2


def test_parallel_head_gate_placement():
    from octave_b200 import network_parallel as npar
    plain = npar.ResnestUnetParallelHead.__new__(npar.ResnestUnetParallelHead)
    assert not any(plain._gate_on(l) for l in range(5))
    for gl, main, par in ((3, [3, 2, 1, 0], [1, 0]), (4, [4, 3, 2, 1, 0], [1, 0]), (0, [0], [0]), (-1, [], [])):
        ag = npar.ResnestUnetParallelHeadAttentionGate.__new__(npar.ResnestUnetParallelHeadAttentionGate)
        object.__setattr__(ag, "gating_level", gl)
        assert [l for l in (4, 3, 2, 1, 0) if ag._gate_on(l)] == main
        assert [l for l in (1, 0) if ag._gate_on(l)] == par


def test_conv_spec_dense_group_merging():
    """Tiny groups are merged into block-diagonal dense groups until a group fills a 32-wide UMMA chunk; a conv the
    tensor-core kernels cannot take reports so instead of being silently mis-run."""
    from octave_b200.ops import ConvSpec
    w = torch.zeros(64, 16, 3, 3)
    s = ConvSpec(w, None, 64, 64, 3, 1, 1, 4)                 # decoder_0 split-attention conv: 4 groups of 16 -> 2 dense
    assert s.dense_groups == 2 and s.tc_ok(torch.bfloat16) and not s.tc_ok(torch.float32)
    s = ConvSpec(torch.zeros(2, 32, 1, 1), None, 32, 2, 1, 1, 0, 1)   # fc head: 2 output channels
    assert s.dense_groups == 0 and not s.tc_ok(torch.bfloat16)
    s = ConvSpec(torch.zeros(32, 3, 3, 3), None, 3, 32, 3, 2, 1, 1)   # stem conv: 3 channels, stride 2
    assert not s.tc_ok(torch.bfloat16)
    s = ConvSpec(torch.zeros(256, 64, 2, 2), None, 256, 64, 2, 2, 0, 1, transposed=True)
    assert s.tc_ok(torch.bfloat16)
