"""octave_b200 — B200-native (sm_100a) kernels behind the OCTAve scribble-supervised training step.

Host code mirrors the reference's Python interface (architectures/*); the arithmetic runs in
hand-written CUDA through the C-ABI of include/octave_b200.h.  No CPU fallback exists.
"""
__version__ = "0.1.0"
