"""Throughput of the C++ flattener alone (Newick text -> kernel-ordered trees) on this host (developer tool)."""
import sys
import time

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, synth  # noqa: E402

for tips, n in ((100, 2000), (500, 4096), (2000, 512)):
    text = "\n".join(synth.tree_set(n, tips, 3000, "mixed"))
    for nthreads in (1, 0):
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            f = FlatTree.many_from_newick(text, 3, "mean", 0, nthreads)
            best = min(best, time.perf_counter() - t0)
            del f
        print("%4d tips x %4d trees, %s: %7.1f ms, %6.1f us/tree, %8.0f trees/s, %.0f MB/s of text"
              % (tips, n, "1 thread " if nthreads == 1 else "all cores", best * 1e3, best * 1e6 / n, n / best,
                 len(text) / best / 1e6), flush=True)
