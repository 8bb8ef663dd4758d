"""Where one sim-vs-obs proposal spends its time on the host (developer tool, not a benchmark)."""
import ctypes
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, PhyloKernel, _lib, get_context, synth  # noqa: E402
from kamphir_b200._lib import check  # noqa: E402


def main():
    tips = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nreps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, normalize="mean")
    obs = synth.kingman_newick(tips, 7)
    sims = [synth.kingman_newick(tips, 2000 + i).encode() for i in range(nreps)]
    pk.score_batch(obs, sims[:2])
    L = _lib.lib()
    T = {}

    def tick(name, t0):
        T.setdefault(name, []).append(time.perf_counter() - t0)

    for _ in range(7):
        t0 = time.perf_counter()
        fo = FlatTree.from_newick(obs)
        fs = FlatTree.many_from_newick_list(sims, 3, "mean", 0, 0)
        tick("flatten (python call)", t0)
        t0 = time.perf_counter()
        n = len(fs)
        k, d, kh = np.empty(n), np.empty(n), np.empty(n)
        ref = ctypes.c_double(float("nan"))
        arr = (ctypes.c_void_p * n)(*[t._h for t in fs])
        p = pk._params()
        tick("marshal", t0)
        t0 = time.perf_counter()
        check(L.pk_score_batch(pk._ctx()._h, fo._h, arr, n, ctypes.byref(p), ctypes.byref(ref), k.ctypes.data,
                               d.ctypes.data, kh.ctypes.data))
        tick("pk_score_batch (C)", t0)
        t0 = time.perf_counter()
        kc = np.array([t.kcoal(fo) for t in fs])
        tick("kcoal x n (python loop)", t0)
        t0 = time.perf_counter()
        res = pk.score_batch(fo, fs, use_kcoal=True)
        tick("score_batch (python, whole)", t0)
        t0 = time.perf_counter()
        del fs, fo
        tick("free trees", t0)
    for name, v in T.items():
        print("%-32s %8.1f us (min of %d)" % (name, 1e6 * min(v), len(v)))
    print("kernel ms", get_context().last_stats()["ms"])


if __name__ == "__main__":
    main()
