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


# ---- tip-state ("labelled") mode: extension without a reference implementation, parity unpinned by the reference ----
def labelled_classes(flat, S):
    """Class id of every internal node of an oracle Flat annotated with tip states: 0 = no tip child, 1 + slot * S + state
    = one tip child, 1 + 2S + state0 * S + state1 = two tip children (states taken mod S: demes 0..S-1 or rcolgem's 1..S)."""
    out = []
    for i in range(flat.n):
        st = flat.state[i]
        tips = [s for s in (0, 1) if flat.cidx[i][s] < 0]
        if not tips:
            out.append(0)
        elif len(tips) == 1:
            out.append(1 + tips[0] * S + st[tips[0]] % S)
        else:
            out.append(1 + 2 * S + (st[0] % S) * S + st[1] % S)
    return out


def labelled_cflat(flat, S):
    """C-oracle input for the tip-state mode: a node's production is its class and a child's production the child's class
    (0 = tip), so the reference's two tests -- equal productions of the nodes, equal productions of the children
    (phyloK2.py:183,191) -- become the labelled ones.  Checked against phylok_oracle.kernel_arrays_labelled and, through
    it, against the recursive definition (tests/test_oracle.py)."""
    cls = labelled_classes(flat, S)
    c = CFlat(flat, classes=np.array(cls, np.uint8) + 1)
    c.cprod = np.ascontiguousarray([[0 if flat.cidx[i][s] < 0 else cls[flat.cidx[i][s]] + 1 for s in (0, 1)]
                                    for i in range(flat.n)], dtype=np.uint8).reshape(-1)
    return c
