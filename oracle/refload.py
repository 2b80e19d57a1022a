"""TEST INFRASTRUCTURE ONLY — import the real reference from /root/reference (when mounted).

The reference imports `kornia.geometry.transform.resize` at module top (segmentor/losses.py:4) and
kornia is not installed; its single call site (losses.py:126) is a nearest-neighbour resize, so a
stand-in module forwarding to torch.nn.functional.interpolate(mode='nearest') is injected.

The repo also ships a package called `architectures` (the drop-in import path); the reference is loaded
with /root/reference first on sys.path and its modules are then removed from sys.modules again so both
can live in one process.  Returns None when /root/reference is absent (the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OCTAVE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "architectures"))


def _install_kornia_shim() -> None:
    if "kornia.geometry.transform" in sys.modules:
        return
    import torch.nn.functional as F

    def resize(input, size, interpolation="nearest", **_):
        assert interpolation == "nearest"
        if tuple(input.shape[-2:]) == tuple(size):
            return input
        return F.interpolate(input, size=size, mode="nearest")

    kornia = types.ModuleType("kornia")
    geometry = types.ModuleType("kornia.geometry")
    transform = types.ModuleType("kornia.geometry.transform")
    transform.resize = resize
    geometry.transform = transform
    kornia.geometry = geometry
    sys.modules["kornia"] = kornia
    sys.modules["kornia.geometry"] = geometry
    sys.modules["kornia.geometry.transform"] = transform


_cache = None


def load():
    """-> namespace with OctaScribbleNet, ResnestUNet, DiscriminatorBlock, the loss classes and blocks."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        return None
    _install_kornia_shim()
    stash = {k: v for k, v in sys.modules.items() if k == "architectures" or k.startswith("architectures.")}
    for k in stash:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib

        ns = types.SimpleNamespace()
        octa = importlib.import_module("architectures.models.octa")
        compose = importlib.import_module("architectures.segmentor.compose")
        seg_losses = importlib.import_module("architectures.segmentor.losses")
        seg_blocks = importlib.import_module("architectures.segmentor.blocks")
        dis_blocks = importlib.import_module("architectures.discriminator.blocks")
        dis_losses = importlib.import_module("architectures.discriminator.losses")
        resnest = importlib.import_module("architectures.extra.resnest")
        assert octa.__file__.startswith(REFERENCE_ROOT), octa.__file__
        ns.OctaScribbleNet = octa.OctaScribbleNet
        ns.ResnestUNet = compose.ResnestUNet
        ns.WeightedPartialCE = seg_losses.WeightedPartialCE
        ns.DiceLoss = seg_losses.DiceLoss
        ns.InterlayerDivergence = seg_losses.InterlayerDivergence
        ns.AdversarialAttentionGate = seg_blocks.AdversarialAttentionGate
        ns.DiscriminatorBlock = dis_blocks.DiscriminatorBlock
        ns.LSDiscriminatorialLoss = dis_losses.LSDiscriminatorialLoss
        ns.LSGeneratorLoss = dis_losses.LSGeneratorLoss
        ns.resnest = resnest
        ns.compose = compose
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "architectures" or k.startswith("architectures.")]:
            del sys.modules[k]
        sys.modules.update(stash)
    _cache = ns
    return ns
