"""Drop-in for pamr.py:115-144 (class PAMR) -- same constructor, same forward signature, same buffer names."""
import torch
import torch.nn as nn

from . import ops


def _shift_kernel(kind):
    # pamr.py:18-36 / :62-77 / :81-98 -- kept only so that state_dicts / buffer names match; the CUDA
    # kernel does the gathers in index arithmetic and never reads these.
    n = 9 if kind == "std" else 8
    w = torch.zeros(n, 1, 3, 3)
    taps = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)]
    if kind == "std":
        taps = [(i // 3, i % 3) for i in range(9)]
    for i, (r, c) in enumerate(taps):
        if kind == "abs":
            w[i, 0, 1, 1] = 1
            w[i, 0, r, c] = -1
        else:
            w[i, 0, r, c] = 1
    return w


class _Aff(nn.Module):
    def __init__(self, kind, dilations):
        super().__init__()
        self.dilations = dilations
        self.register_buffer("kernel", _shift_kernel(kind))


class PAMR(nn.Module):
    def __init__(self, num_iter=1, dilations=[1]):
        super().__init__()
        self.num_iter = num_iter
        self.dilations = list(dilations)
        self.aff_x = _Aff("abs", self.dilations)
        self.aff_m = _Aff("copy", self.dilations)
        self.aff_std = _Aff("std", self.dilations)

    def forward(self, x, mask):
        """x [B,K,H,W] image, mask [B,C,h,w] -> refined mask [B,C,H,W] (pamr.py:125-144)."""
        return ops.pamr_forward(x, mask, self.dilations, self.num_iter)
