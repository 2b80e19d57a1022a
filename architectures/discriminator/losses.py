from octave_b200.losses import LSDiscriminatorialLoss, LSGeneratorLoss  # noqa: F401
