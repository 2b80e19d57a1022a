"""Run-time switches of the host layer (no effect on numerics of a given mode)."""

# InterlayerDivergence raises Exception('Divergence is NaN') like the reference (losses.py:140-142).
# The check is a device->host read; the benchmark and FusedSegmentorLoss read the flag lazily instead.
nan_check = True

# Arithmetic/storage type of the network's activations: 'bf16' (tensor-core path) or 'fp32'.
compute_dtype = "bf16"


# Inference (eval mode under torch.no_grad()): fold every BatchNorm that directly follows a convolution into that
# convolution's weights / epilogue.  False keeps the separate BN pass (same results to rounding).
fold_bn_inference = True


def set_compute_dtype(name: str) -> None:
    global compute_dtype
    if name not in ("bf16", "fp32"):
        raise ValueError("compute dtype must be 'bf16' or 'fp32'")
    compute_dtype = name
