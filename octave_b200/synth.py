"""Seeded synthetic OCTA-like inputs: en face image with vessel-like ridges, sparse
one-hot scribbles (~3 % per class, rest unlabelled = all-zero rows), and a real-mask pyramid for D."""
from __future__ import annotations

import math

import torch


def _ridges(gen: torch.Generator, B: int, H: int, W: int, n: int = 40) -> torch.Tensor:
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    out = torch.zeros(B, H, W)
    for b in range(B):
        acc = torch.zeros(H, W)
        for _ in range(n):
            r = torch.rand(6, generator=gen)
            theta = float(r[0]) * math.pi
            off = float(r[1]) * max(H, W)
            amp = 4.0 + 20.0 * float(r[2])
            freq = 0.01 + 0.04 * float(r[3])
            width = 0.8 + 2.5 * float(r[4])
            d = (xx * math.cos(theta) + yy * math.sin(theta)) - off - amp * torch.sin(freq * (yy * math.cos(theta) - xx * math.sin(theta)) + 6.28 * float(r[5]))
            acc = torch.maximum(acc, torch.exp(-(d / width) ** 2))
        out[b] = acc
    return out


def octa_batch(B: int, H: int, W: int, seed: int = 0, n_ridges: int = 40):
    """-> x [B,3,H,W] in [0,1], scribble ys [B,2,H,W] one-hot/zero, vessel ground truth [B,H,W] bool."""
    gen = torch.Generator().manual_seed(seed)
    ridge = _ridges(gen, B, H, W, n_ridges)
    bg = torch.rand(B, H, W, generator=gen) * 0.3
    img = torch.clamp(bg + ridge * (0.5 + 0.5 * torch.rand(B, 1, 1, generator=gen)), 0, 1)
    x = img.unsqueeze(1).repeat(1, 3, 1, 1).contiguous()
    vessel = ridge > 0.5
    pick = torch.rand(B, H, W, generator=gen)
    fg = vessel & (ridge > 0.9) & (pick < 0.35)
    bgs = (~vessel) & (ridge < 0.05) & (pick < 0.04)
    ys = torch.stack([bgs.float(), fg.float()], dim=1).contiguous()
    return x, ys, vessel


def mask_pyramid(B: int, H: int, W: int, levels: int = 5, seed: int = 1, n_ridges: int = 40):
    """Unpaired 'real' one-hot masks, nearest-downsampled to H/2^k (input of the mask discriminator)."""
    gen = torch.Generator().manual_seed(seed)
    vessel = _ridges(gen, B, H, W, n_ridges) > 0.5
    full = torch.stack([(~vessel).float(), vessel.float()], dim=1)
    return [full[:, :, :: 2 ** k, :: 2 ** k].contiguous() for k in range(levels)]


def prob_maps(B: int, C: int, H: int, W: int, levels: int, seed: int = 0):
    """Softmax probability pyramid [B,C,H>>k,W>>k] from random logits."""
    gen = torch.Generator().manual_seed(seed)
    return [torch.softmax(2.0 * torch.randn(B, C, H >> k, W >> k, generator=gen), dim=1) for k in range(levels)]
