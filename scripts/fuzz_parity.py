"""Randomised parity sweep of the CUDA path against the oracles (developer tool; the fixed cases live in tests/).
    python scripts/fuzz_parity.py [rounds] [seed]
Each round draws a tree set with ragged sizes (2..900 tips, Kingman / Yule / caterpillar / balanced, optional tip
states), kernel parameters, normalisation, precision and a launch configuration (CTA size, C storage), evaluates random
pairs and compares with the C oracle (unlabelled) or the Python extension oracle (labelled)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, PhyloKernel, get_context, synth  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import phylok_oracle as po  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    worst = {64: 0.0, 32: 0.0}
    checked = 0
    t0 = time.time()
    for rd in range(rounds):
        S = int(rng.choice([0, 0, 0, 2, 3]))
        big = S == 0 and rng.random() < 0.3
        ntrees = int(rng.integers(3, 9))
        nwk = []
        for k in range(ntrees):
            hi = 900 if big else (160 if S else 320)
            n = int(rng.integers(2, hi))
            kind = rng.choice(["kingman", "yule", "cat", "bal"], p=[0.45, 0.35, 0.1, 0.1])
            sd = int(rng.integers(1, 10 ** 6))
            if kind == "kingman":
                nwk.append(synth.kingman_newick(n, sd, states=S))
            elif kind == "yule":
                nwk.append(synth.yule_newick(n, sd, states=S))
            elif kind == "cat" and not S:
                nwk.append(synth.caterpillar_newick(min(n, 200), sd))
            elif kind == "bal" and not S:
                nwk.append(synth.balanced_newick(int(rng.integers(1, 8)), sd))
            else:
                nwk.append(synth.kingman_newick(n, sd, states=S))
        norm = str(rng.choice(["mean", "median", "none"]))
        lam = float(rng.choice([0.05, 0.1, 0.2, 0.5, 1.0]))
        tau = float(rng.choice([0.3, 1.0, 2.0, 10.0]))
        sig = float(rng.choice([0.5, 1.0, 2.0]))
        prec = int(rng.choice([64, 64, 32]))
        threads = int(rng.choice([0, 0, 64, 128, 256]))
        storage = int(rng.choice([0, 0, 1, 2])) if not big else 0
        # where the tree parts are staged: automatic, both parts in global memory, COLUMN part in global memory (unlabelled)
        stage = int(rng.choice([0, 0, 1, 2]))
        if stage == 2 and (S or threads):
            stage = 1
        get_context().configure(threads, 0, storage)
        get_context().set_option("unstaged", stage)
        flats = [FlatTree.from_newick(s, 3, norm, S) for s in nwk]
        pairs = [(int(a), int(b)) for a, b in rng.integers(0, ntrees, size=(12, 2))] + [(0, 0)]
        pk = PhyloKernel(decayFactor=lam, gaussFactor=tau, sigma=sig, precision=prec, n_states=S, normalize=norm)
        try:
            got = pk.kernel_batch(flats, flats, pairs=pairs)
        except Exception as exc:          # smem storage may not fit: that is an error the library reports, not a mismatch
            if storage == 2:
                continue
            raise
        if S:
            of = [po.prep_kamphir(po.read_newick(s), norm, po.state_from_name) for s in nwk]
            want = [po.kernel_arrays_labelled(of[a], of[b], lam, tau, sig) for a, b in pairs]
        else:
            of = [c_oracle.CFlat(po.prepare(po.read_newick(s), "kamphir:" + norm)) for s in nwk]
            want = [c_oracle.kernel(of[a], of[b], lam, tau, sig) for a, b in pairs]
        got, want = np.asarray(got), np.asarray(want)
        # cells overflow beyond 3e38 in fp32 (1.8e308 in fp64, in the oracle too) with decayFactor near 1 on deep, similar trees
        keep = want < 1e300      # fp32 overflow (beyond 3e38) is re-evaluated in fp64 by the library itself
        got, want = got[keep], want[keep]
        if not len(want):
            continue
        if not np.all(np.isfinite(got)):
            print("NON-FINITE result in round %d" % rd)
            sys.exit(1)
        err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))
        worst[prec] = max(worst[prec], err)
        checked += len(pairs)
        tol = 1e-9 if prec == 64 else 1e-5
        if prec == 32 and np.any(want > 1e30):      # near the fp32 range the cells lose their last bits before they overflow
            tol = 1e-4
        if err > tol:
            print("MISMATCH round %d: S=%d norm=%s lam=%g tau=%g sig=%g prec=%d threads=%d storage=%d stage=%d err=%.3g sizes=%s"
                  % (rd, S, norm, lam, tau, sig, prec, threads, storage, stage, err, [f.n_internal + 1 for f in flats]))
            sys.exit(1)
    get_context().configure(0, 0, 0)
    get_context().set_option("unstaged", 0)
    print("fuzz ok: %d rounds, %d pairs, worst relative error fp64 %.3g fp32 %.3g, %.1f s"
          % (rounds, checked, worst[64], worst[32], time.time() - t0))


if __name__ == "__main__":
    main()
