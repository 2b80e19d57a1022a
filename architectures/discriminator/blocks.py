from octave_b200.discriminator import DiscriminatorBlock, InstanceNoise, LabelNoise  # noqa: F401
