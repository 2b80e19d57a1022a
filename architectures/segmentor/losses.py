from octave_b200.losses import DiceLoss, InterlayerDivergence, WeightedPartialCE  # noqa: F401
