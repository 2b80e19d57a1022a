from octave_b200.network import ResnestUNet  # noqa: F401
