"""Drop-in for the SWIG module wrapper/bilateralfilter/bilateralfilter.py (bilateralfilter.i:21-25).

`bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy)` takes 1-D contiguous float32 numpy
arrays exactly like the SWIG export (IN_ARRAY1 / INPLACE_ARRAY1 typemaps) and writes `outs` in place.  It
calls the host-buffer C entry point `bilateralfilter_batch_b200`, which stages through the GPU.
"""
import ctypes

import numpy as np

from . import _lib


def _check(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != 1 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name}: expected a 1-D contiguous float32 numpy array (as the SWIG typemap requires)")


def bilateralfilter_batch(images, ins, outs, N, K, H, W, sigmargb, sigmaxy):
    _check(images, "images"); _check(ins, "ins"); _check(outs, "outs")
    if not outs.flags["WRITEABLE"]:
        raise TypeError("outs must be writable (INPLACE_ARRAY1)")
    lib = _lib.lib()
    lib.acr_last_error_string()  # make sure the library is initialised
    lib.bilateralfilter_batch_b200(images.ctypes.data_as(ctypes.c_void_p), images.size,
                                   ins.ctypes.data_as(ctypes.c_void_p), ins.size,
                                   outs.ctypes.data_as(ctypes.c_void_p), outs.size,
                                   int(N), int(K), int(H), int(W), float(sigmargb), float(sigmaxy))
    err = _lib.last_error()
    if err:
        raise RuntimeError(err)


def bilateralfilter(image, in_, out, H, W, sigmargb, sigmaxy):
    """Single-image form (bilateralfilter.hpp:11): K is inferred from len(in)/(H*W)."""
    K = in_.size // (H * W)
    bilateralfilter_batch(image, in_, out, 1, K, H, W, sigmargb, sigmaxy)
