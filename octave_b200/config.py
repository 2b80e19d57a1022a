"""Run-time switches of the host layer (no effect on numerics of a given mode)."""
import os

# InterlayerDivergence raises Exception('Divergence is NaN') like the reference (losses.py:140-142).
# The check is a device->host read; the benchmark and FusedSegmentorLoss read the flag lazily instead.
nan_check = True

# Arithmetic/storage type of the network's activations: 'bf16' (tensor-core path) or 'fp32'.
compute_dtype = "bf16"


# Inference (eval mode under torch.no_grad()): fold every BatchNorm that directly follows a convolution into that
# convolution's weights / epilogue.  False keeps the separate BN pass (same results to rounding).
fold_bn_inference = True


# Backward pass: launch the weight-gradient kernels on a second stream.  wgrad(L) depends only on the saved input of conv L
# and on dz_L, so it can run beside the rest of the chain (dgrad(L), then the HBM-bound BatchNorm backward of layer L-1):
# tensor-pipe-bound and HBM-bound kernels then share the SMs instead of taking turns.  Joined before the gradients are used.
overlap_wgrad = True
# The attention branch of SplAtConv2d (fc1 -> bn1 -> relu -> fc2 -> r-softmax) when the batch fits one 32-row slab
# (octave_attn_fused_supported).  Forward: two launches instead of four (-0.35 ms per c2 step).  Backward: the single-launch
# form (r-softmax backward + fc2 data gradient + bn1/relu backward) is bit-identical too but its 1-8 blocks walk all 2C
# outputs serially, +0.7 ms per step against the split-K kernels: off unless OCTAVE_FUSE_ATTN=2.  OCTAVE_FUSE_ATTN=0: all separate.
fuse_attention_branch = os.environ.get("OCTAVE_FUSE_ATTN", "1") != "0"
fuse_attention_branch_bwd = os.environ.get("OCTAVE_FUSE_ATTN", "1") == "2"


def set_compute_dtype(name: str) -> None:
    global compute_dtype
    if name not in ("bf16", "fp32"):
        raise ValueError("compute dtype must be 'bf16' or 'fp32'")
    compute_dtype = name


def set_deterministic(on: bool) -> None:
    """Bit-reproducible mode: reductions that are normally split across CTAs and merged with fp32 atomics (split-K weight
    gradients, K-split linears) run with one writer per output element (include/octave_b200.h: octave_set_deterministic).
    The forward pass (BatchNorm statistics in fp64, fixed-order global average pool, single-pass loss) is reproducible in
    either mode."""
    from . import _lib
    _lib.lib.octave_set_deterministic(1 if on else 0)
