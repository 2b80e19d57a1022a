from octave_b200.synth import *  # noqa: F401,F403
from octave_b200.synth import mask_pyramid, octa_batch, prob_maps  # noqa: F401
