from octave_b200.network import ResnestUNet  # noqa: F401
from octave_b200.network_parallel import ResnestUnetParallelHead, ResnestUnetParallelHeadAttentionGate  # noqa: F401
