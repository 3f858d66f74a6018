"""Seeded synthetic inputs of the shapes BASELINE.json names.  CPU generators only;
no dataset, no checkpoint -- used by tests/, bench.py and __graft_entry__.smoke()."""
import math

import torch


def images(B, S, seed=0):
    """Normalised-space images [B,3,S,S] (torch.randn, generator seed)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, S, S, generator=g)


def labels(B, C, seed=0, present=None):
    """Multi-hot labels [B,C] with at least one positive per row; `present` forces an explicit class set."""
    if present is not None:
        y = torch.zeros(B, C)
        y[:, list(present)] = 1.0
        return y
    g = torch.Generator().manual_seed(seed + 1)
    y = (torch.rand(B, C, generator=g) > 0.8).float()
    for b in range(B):
        if y[b].sum() == 0:
            y[b, b % C] = 1.0
    return y


def smooth_rgb(B, H, W, seed=0, noise=5.0):
    """0..255 images: sum of 8 low-frequency sinusoids per channel + N(0, noise), clipped (PAMR / bilateral)."""
    g = torch.Generator().manual_seed(seed + 2)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    out = torch.zeros(B, 3, H, W)
    for b in range(B):
        for c in range(3):
            acc = torch.zeros(H, W)
            for _ in range(8):
                fx, fy = (torch.rand(2, generator=g) * 3.0 + 0.25).tolist()
                ph = float(torch.rand(1, generator=g)) * 2 * math.pi
                amp = float(torch.rand(1, generator=g)) * 24 + 8
                acc += amp * torch.sin(2 * math.pi * (fx * xx / W + fy * yy / H) + ph)
            out[b, c] = 128 + acc
    out += noise * torch.randn(B, 3, H, W, generator=g)
    return out.clamp_(0, 255)


def probabilities(B, C, H, W, seed=0):
    g = torch.Generator().manual_seed(seed + 3)
    return torch.softmax(torch.randn(B, C, H, W, generator=g), dim=1)
