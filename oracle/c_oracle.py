"""ctypes front end of oracle/libpk_oracle.so (C restatement of phyloK2.py:158-209).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C oracle in place (gcc, seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpk_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        P = ctypes.c_void_p
        L.pk_oracle_kernel.restype = ctypes.c_double
        L.pk_oracle_kernel.argtypes = [ctypes.c_int32, P, P, P, P, ctypes.c_int32, P, P, P, P,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_double, P]
        _LIB = L
    return _LIB


class CFlat(object):
    """numpy copy of oracle.phylok_oracle.Flat (post-order arrays)."""

    def __init__(self, flat, classes=None):
        self.n = flat.n
        self.prod = np.ascontiguousarray(classes if classes is not None else flat.prod, dtype=np.uint8)
        self.cprod = np.ascontiguousarray(flat.cprod, dtype=np.uint8).reshape(-1)
        self.cidx = np.ascontiguousarray(flat.cidx, dtype=np.int32).reshape(-1)
        self.bl = np.ascontiguousarray(flat.bl, dtype=np.float64).reshape(-1)


def kernel(a, b, decayFactor=0.1, gaussFactor=1., sigma=1., counts=False):
    """K(a, b) for two CFlat objects; with counts=True also (node pairs, M, R)."""
    cnt = np.zeros(3, dtype=np.int64)
    k = lib().pk_oracle_kernel(a.n, a.prod.ctypes.data, a.cprod.ctypes.data, a.cidx.ctypes.data, a.bl.ctypes.data,
                               b.n, b.prod.ctypes.data, b.cprod.ctypes.data, b.cidx.ctypes.data, b.bl.ctypes.data,
                               float(decayFactor), float(gaussFactor), float(sigma), cnt.ctypes.data)
    if counts:
        return k, tuple(int(x) for x in cnt)
    return k
