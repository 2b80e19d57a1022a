from octave_b200.model import OctaScribbleNet  # noqa: F401
