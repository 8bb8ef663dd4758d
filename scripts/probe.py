"""Quick device-side timing probe (not the benchmark): Gram part over n trees for several launch configurations."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, Forest, _lib, get_context, synth  # noqa: E402
from kamphir_b200._lib import check, pk_params  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    tips = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    prec_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64, 32]
    cfgs = sys.argv[4] if len(sys.argv) > 4 else "0:0:0,512:1:1,512:2:1,1024:1:1,256:2:1,256:4:1"
    states = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # tip states (BASELINE.json configs[2]): labelled productions
    t0 = time.time()
    nwk = synth.tree_set(n, tips, 3000, "mixed", states, [1.0 / states] * states if states else None)
    if os.environ.get("PK_PROBE_SAME"):          # n copies of one tree: every pair has the forest's maximal group sizes
        nwk = [nwk[0]] * n
    flats = FlatTree.many_from_newick("\n".join(nwk), 3, "mean", states)
    print("gen+flatten %.2fs" % (time.time() - t0), flush=True)
    ctx = get_context(0)
    L = _lib.lib()
    d_out, d_diag = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.pk_device_alloc(ctx._h, 8 * n * n, ctypes.byref(d_out)))
    check(L.pk_device_alloc(ctx._h, 8 * n, ctypes.byref(d_diag)))
    for prec in prec_list:
        forest = Forest(ctx, flats, prec)
        st = forest.gram_stats(32)
        w = 8 if prec == 64 else 4
        node_sum = 2 * (tips - 1) * st["dps"]
        alg = w * (st["cells"] + st["child_reads"]) + 32 * node_sum + 8 * st["dps"]
        p = pk_params(0.2, 2.0, 1.0, prec, 0)
        for cfg in cfgs.split(","):
            th, ctas, force = [int(x) for x in cfg.split(":")]
            ctx.configure(th, ctas, force)
            best = 1e30
            for rep in range(3):
                check(L.pk_forest_gram_part(ctx._h, forest._h, ctypes.byref(p), 1, 32, 0, 1, d_out, n, d_diag))
                s = ctx.last_stats()
                best = min(best, s["ms"])
            print("fp%d cfg=%-10s path=%s grid=%d thr=%d smem=%d  %.3f ms  %.3f MDP/s  %.2f Gcells/s  alg %.1f GB/s (%.3f of 6536)"
                  % (prec, cfg, s["path"], s["grid"], s["threads"], s["smem_bytes"], best, st["dps"] / best / 1e3,
                     st["cells"] / best / 1e6, alg / best / 1e6, alg / best / 1e6 / 6536.4), flush=True)
        forest.close()


if __name__ == "__main__":
    main()
