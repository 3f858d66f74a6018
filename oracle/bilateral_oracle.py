"""ctypes doors onto the two CPU checkers of the bilateral filter.  TEST INFRASTRUCTURE ONLY.

  oracle_bilateral : oracle/permuto_oracle.c (this repo's C restatement; always available after
                     `make -C oracle`)
  ref_bilateral    : oracle/_ref/libbilateral_ref.so = the reference's own C++
                     (wrapper/bilateralfilter/{bilateralfilter,permutohedral}.cpp) compiled in the build
                     container; travels to the GPU box as a prebuilt file, absent in a fresh clone.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libpermuto_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libbilateral_ref.so")
_ARGS = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def _call(fn, images, ins, sigmargb, sigmaxy):
    images = np.ascontiguousarray(images, dtype=np.float32)
    ins = np.ascontiguousarray(ins, dtype=np.float32)
    N, K, H, W = ins.shape
    assert images.shape == (N, 3, H, W)
    outs = np.zeros_like(ins)
    fn(images.ctypes.data, images.size, ins.ctypes.data, ins.size, outs.ctypes.data, outs.size,
       N, K, H, W, float(sigmargb), float(sigmaxy))
    return outs


def oracle_bilateral(images, ins, sigmargb, sigmaxy):
    if not os.path.exists(_ORACLE_SO):
        build()
    lib = ctypes.CDLL(_ORACLE_SO)
    lib.permuto_oracle_batch.argtypes = _ARGS
    lib.permuto_oracle_batch.restype = None
    return _call(lib.permuto_oracle_batch, images, ins, sigmargb, sigmaxy)


def have_ref():
    return os.path.exists(_REF_SO)


def ref_bilateral(images, ins, sigmargb, sigmaxy):
    lib = ctypes.CDLL(_REF_SO)
    lib.ref_bilateralfilter_batch.argtypes = _ARGS
    lib.ref_bilateralfilter_batch.restype = None
    return _call(lib.ref_bilateralfilter_batch, images, ins, sigmargb, sigmaxy)
