from octave_b200.discriminator import rand_uniform  # noqa: F401
