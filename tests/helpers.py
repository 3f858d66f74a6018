import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative error' of the north star: scale-normalised max error)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def t2n(t):
    return t.detach().float().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
