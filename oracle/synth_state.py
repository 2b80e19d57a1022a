"""TEST / BASELINE INFRASTRUCTURE ONLY — a seeded OctaScribbleNet state dict built from the committed shape table
(oracle/state_shapes.json: the 682+ keys / shapes of the reference's state_dict, models/octa.py:44-57) with plain torch,
so that the CPU reference arm of bench.py needs neither /root/reference nor this repo's CUDA library in its process.
Initialisation follows the reference's rules in spirit (resnest.py:368-374: conv ~ N(0, sqrt(2/n)), BatchNorm weight 1 /
bias 0); the values only feed timing runs."""
from __future__ import annotations

import json
import math
import os

import torch


def seeded_state(H: int, W: int, seed: int = 0):
    tab = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_shapes.json")))
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, dtype in tab:
        shape = [H // 32 if s == "H/32" else W // 32 if s == "W/32" else s for s in shape]
        if dtype == "int64":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
        elif key.endswith("running_var"):
            sd[key] = torch.ones(shape)
        elif key.endswith("running_mean"):
            sd[key] = torch.zeros(shape)
        elif key.endswith(("_u", "_v")):
            v = torch.randn(shape, generator=g)
            sd[key] = v / v.norm()
        elif len(shape) == 4:
            n = shape[2] * shape[3] * shape[0]
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / n)
        elif len(shape) == 2:
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * (2.0 / math.sqrt(shape[1]))
        elif key.endswith("weight"):
            sd[key] = torch.ones(shape)          # BatchNorm scale
        else:
            sd[key] = torch.zeros(shape)         # biases / BatchNorm shift
    return sd
