"""Host-side mirror vs the real reference (CPU, needs /root/reference): identical state_dict keys/shapes, identical
seeded initialisation and RNG consumption, state_dict round trip both ways, constructor/attribute surface."""
import inspect

import pytest
import torch

from oracle import refload

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("kw", [{}, {"instance_noise": False, "label_noise": False, "discriminator_depth": 3}])
def test_state_dict_and_seeded_init_match_reference(kw):
    ref = refload.load()
    from architectures.models.octa import OctaScribbleNet
    shape = (torch.Size((2, 3, 304, 304)), torch.Size((2, 2, 304, 304)))
    torch.manual_seed(0)
    r = ref.OctaScribbleNet(*shape, True, False, **kw)
    ra = torch.rand(4)
    torch.manual_seed(0)
    m = OctaScribbleNet(*shape, True, False, **kw)
    rb = torch.rand(4)
    a, b = r.state_dict(), m.state_dict()
    assert list(a.keys()) == list(b.keys()) and len(a) in (682, 672, 662, 652) or len(a) == len(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    assert torch.equal(ra, rb), "construction must consume the global RNG exactly like the reference"
    m.load_state_dict(r.state_dict())
    r.load_state_dict(m.state_dict())
    for attr in ("segmentor", "discriminator", "supervised_loss", "discriminatorial_loss", "generator_loss", "is_train"):
        assert hasattr(m, attr)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1))


def test_signatures_match_reference():
    ref = refload.load()
    from octave_b200 import discriminator, losses, model, network
    pairs = [(ref.OctaScribbleNet.__init__, model.OctaScribbleNet.__init__),
             (ref.ResnestUNet.__init__, network.ResnestUNet.__init__),
             (ref.DiscriminatorBlock.__init__, discriminator.DiscriminatorBlock.__init__),
             (ref.WeightedPartialCE.__init__, losses.WeightedPartialCE.__init__),
             (ref.WeightedPartialCE.forward, losses.WeightedPartialCE.forward),
             (ref.DiceLoss.forward, losses.DiceLoss.forward),
             (ref.InterlayerDivergence.__init__, losses.InterlayerDivergence.__init__),
             (ref.InterlayerDivergence.forward, losses.InterlayerDivergence.forward),
             (ref.LSDiscriminatorialLoss.forward, losses.LSDiscriminatorialLoss.forward),
             (ref.LSGeneratorLoss.forward, losses.LSGeneratorLoss.forward)]
    for a, b in pairs:
        pa, pb = inspect.signature(a).parameters, inspect.signature(b).parameters
        assert list(pa) == list(pb), (a, list(pa), list(pb))
        for k in pa:
            assert pa[k].default == pb[k].default or pa[k].default is inspect._empty, (a, k)


def test_no_discriminator_attribute_when_depth_is_zero():
    from architectures.models.octa import OctaScribbleNet
    m = OctaScribbleNet(torch.Size((2, 3, 64, 64)), torch.Size((2, 2, 64, 64)), True, False, discriminator_depth=0)
    assert not hasattr(m, "discriminator")   # models/octa.py:46


@pytest.mark.parametrize("gated", [False, True])
def test_parallel_head_siblings_match_reference(gated):
    """ResnestUnetParallelHead / ...AttentionGate (compose.py:233-527): same keys, shapes, seeded init, RNG consumption
    and constructor signature as the reference."""
    ref = refload.load()
    from architectures.segmentor import compose
    name = "ResnestUnetParallelHeadAttentionGate" if gated else "ResnestUnetParallelHead"
    rc, mc = getattr(ref.compose, name), getattr(compose, name)
    torch.manual_seed(0)
    r = rc(2, False)
    ra = torch.rand(4)
    torch.manual_seed(0)
    m = mc(2, False)
    rb = torch.rand(4)
    a, b = r.state_dict(), m.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    assert torch.equal(ra, rb)
    m.load_state_dict(r.state_dict())
    r.load_state_dict(m.state_dict())
    pa, pb = inspect.signature(rc.__init__).parameters, inspect.signature(mc.__init__).parameters
    assert list(pa) == list(pb)
    for k in pa:
        assert pa[k].default == pb[k].default or pa[k].default is inspect._empty
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))
