#!/usr/bin/env python
"""Benchmark of the phyloK2 tree-kernel hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference's own phyloK2.py on the host cores, rank 0 only

Headline workload (BASELINE.json configs[3], the configuration the metric is quoted on): the all-pairs normalised Gram
matrix of 4096 synthetic 500-tip trees (2048 Kingman + 2048 Yule), lambda=0.2, tau=2.0, sigma=1, kamphir-style prep.  One
"step" = one full Gram matrix = N(N+1)/2 tree-pair DPs, sharded over the ranks by tiles (strong scaling: the matrix is
fixed, more GPUs share it).  `value` = pair evaluations per second with the packed forest already resident in HBM;
`e2e` = the same through host buffers (packed-tree upload H2D, result matrix D2H inside the timed region), median of 5.

At N = 1 the line also carries, measured outside the timed Gram loop:
  `configs`       the other four BASELINE.json configurations ([0] tests/test.tree vs 100 x 100-tip trees, [1] 100 x 200 tips,
                  [2] 500-tip multi-type trees with tip-state matching on and off, [4] 256 x 2000 tips), each through
                  PhyloKernel.score_batch (Newick bytes in -> scores out) with its device-only kernel time, roofline,
                  CPU sample and a parity verdict against the oracle;
  `kernel_call`   latency of the call kamphir.py itself makes, PhyloKernel.kernel(tree object, tree object);
  `cpu_baseline`  the reference's own phyloK2.PhyloKernel.kernel (oracle/_ref/phyloK2.py, unmodified) on one host core.
The run exits with status 3 when any parity check fails (Gram gate, reference literals, configs).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
import multiprocessing as mp
import os
import re
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

LAM, TAU, SIGMA = 0.2, 2.0, 1.0
BASE_SEED = 3000            # seed = config index * 1000 + tree index (config index 3)
T1 = "( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;"      # reference tests/unittests.py:87
T2 = "( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;"    # reference tests/unittests.py:88


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trees", type=int, default=4096)
    ap.add_argument("--tips", type=int, default=500)
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--tile", type=int, default=32)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "allreduce"])
    ap.add_argument("--cpu-sample", type=int, default=160, help="DPs in the single-process cpu_baseline sample (~20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0],[1],[2],[4] block (N = 1 only anyway)")
    ap.add_argument("--parity-sample", type=int, default=256)
    ap.add_argument("--e2e-reps", type=int, default=5)
    ap.add_argument("--corrupt", action="store_true", help="self-test of the gate: spoil one fetched entry, expect exit 3")
    return ap.parse_args()


def make_trees(n_trees, n_tips):
    from kamphir_b200 import synth
    return synth.tree_set(n_trees, n_tips, BASE_SEED, "mixed")


def median(xs):
    xs = sorted(xs)
    m = len(xs) // 2
    return xs[m] if len(xs) % 2 else 0.5 * (xs[m - 1] + xs[m])


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------------
# CPU arms.  "reference" = the reference's own phyloK2.py (oracle/_ref, placed by oracle/Makefile, unmodified) under the
# image's CPython; "port" = the pure-Python restatement (only when oracle/_ref did not travel).  kernel() calls only, trees
# prepared beforehand the way kamphir.py:279-282 prepares them.
# ---------------------------------------------------------------------------------------------------
_CPU = {}


def cpu_kind():
    from oracle import ref_runner
    return "reference" if ref_runner.available() else "port"


def _cpu_init(newicks, lam=LAM, tau=TAU, sigma=SIGMA):
    kind = cpu_kind()
    if kind == "reference":
        from oracle import ref_runner
        pk = ref_runner.make_kernel(lam, tau, sigma, "mean")
        trees = [ref_runner.prepare(pk, s, "mean") for s in newicks]
        _CPU.update(kind=kind, trees=trees, fn=lambda a, b: pk.kernel(trees[a], trees[b]))
    else:
        from oracle import phylok_oracle as po
        flats = [po.prep_kamphir(po.read_newick(s), "mean") for s in newicks]
        _CPU.update(kind=kind, trees=flats, fn=lambda a, b: po.kernel_arrays(flats[a], flats[b], lam, tau, sigma))


def _cpu_dp(pair):
    t0 = time.perf_counter()
    k = _CPU["fn"](pair[0], pair[1])
    return k, time.perf_counter() - t0


def cpu_note():
    if cpu_kind() == "reference":
        return ("the reference's own phyloK2.PhyloKernel.kernel (oracle/_ref/phyloK2.py, unmodified copy of "
                "/root/reference/phyloK2.py placed by oracle/Makefile) under CPython %s with the Bio.Phylo stand-in of "
                "oracle/biophylo_shim; the reference's Python 2.7 + Biopython are absent from the image"
                % sys.version.split()[0])
    return ("pure-Python port of phyloK2.PhyloKernel.kernel (oracle/_ref/phyloK2.py did not travel to this box) under "
            "CPython %s" % sys.version.split()[0])


def sample_pairs(n_trees, count, seed=7):
    import numpy as np
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n_trees, size=count)
    j = rng.integers(0, n_trees, size=count)
    return [(int(min(a, b)), int(max(a, b))) for a, b in zip(i, j)]


def cpu_baseline_single(newicks, pairs):
    """Single process, the headline denominator ('single-process CPU phyloK2'): kernel() calls only, prep excluded."""
    used = sorted(set(x for p in pairs for x in p))
    remap = {t: k for k, t in enumerate(used)}
    _cpu_init([newicks[t] for t in used])
    t0 = time.perf_counter()
    ks = [_cpu_dp((remap[a], remap[b]))[0] for a, b in pairs]
    dt = time.perf_counter() - t0
    return len(pairs) / dt, ks, dt


def port_rate(newicks, pairs):
    """The pure-Python port (round 1's baseline), kept as a secondary figure next to the real reference."""
    from oracle import phylok_oracle as po
    used = sorted(set(x for p in pairs for x in p))
    flats = {t: po.prep_kamphir(po.read_newick(newicks[t]), "mean") for t in used}
    t0 = time.perf_counter()
    for a, b in pairs:
        po.kernel_arrays(flats[a], flats[b], LAM, TAU, SIGMA)
    return len(pairs) / (time.perf_counter() - t0)


def c_oracle_rate(newicks, pairs):
    from oracle import c_oracle
    from oracle import phylok_oracle as po
    used = sorted(set(x for p in pairs for x in p))
    flats = {t: c_oracle.CFlat(po.prep_kamphir(po.read_newick(newicks[t]), "mean")) for t in used}
    t0 = time.perf_counter()
    ks = [c_oracle.kernel(flats[a], flats[b], LAM, TAU, SIGMA) for a, b in pairs]
    dt = time.perf_counter() - t0
    return len(pairs) / dt, ks


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (phyloK2.PhyloKernel.kernel, oracle/_ref) on
    all host cores, one process per core, each step a bounded sample of the Gram workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    per_core = 8 if kind == "reference" else 32      # ~1 s of work per step on every core (0.12 s per DP for the reference)
    per_step = max(per_core * cores, 64)
    newicks = make_trees(min(args.trees, 256), args.tips)
    n = len(newicks)
    steps = args.warmup + args.steps
    pairs = sample_pairs(n, per_step * steps, seed=11)
    with mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(newicks,)) as pool:
        times = []
        for s in range(steps):
            chunk = pairs[s * per_step:(s + 1) * per_step]
            t0 = time.perf_counter()
            pool.map(_cpu_dp, chunk, chunksize=1)
            times.append(time.perf_counter() - t0)
    timed = times[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = per_step / (ms / 1e3)
    sample = "%d random (i<=j) pairs of the first %d trees per step, %d processes" % (per_step, n, cores)
    line = {
        "impl": "reference", "metric": "normalised-kernel pair-evals/s (Gram matrix, 500-tip trees)", "value": value,
        "unit": "pair-evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": cores, "kind": kind, "sample": sample,
                         "note": cpu_note()},
        "e2e": {"value": value, "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, **extra):
    cfg = {"workload": "all-pairs normalised Gram matrix of %d synthetic %d-tip trees (half Kingman, half Yule), "
                       "BASELINE.json configs[3]" % (args.trees, args.tips),
           "n_trees": args.trees, "n_tips": args.tips, "decayFactor": LAM, "gaussFactor": TAU, "sigma": SIGMA,
           "normalize": "mean", "precision": args.precision, "tile": args.tile}
    cfg.update(extra)
    return cfg


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback 6650 (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# the reference's own unit-test values through the GPU path, asserted at start-up
# ---------------------------------------------------------------------------------------------------
def check_literals(precision):
    from kamphir_b200 import PhyloKernel
    pk = PhyloKernel(decayFactor=0.5, gaussFactor=1, precision=precision)
    tol = 1e-9 if precision == 64 else 1e-5
    want1 = 1.125 * (1 + math.exp(-0.0625))                              # tests/unittests.py:95
    k1 = pk.kernel(T1, T2)
    text = open(os.path.join(REPO, "tests", "golden", "trees", "test.tree")).read()
    trees = [t.strip() + ";" for t in text.split(";") if t.strip()]
    k2 = pk.kernel(trees[0], trees[1])                                  # tests/unittests.py:101 (7 places)
    ok2 = round(k2 - 3042.58677467, 7) == 0 if precision == 64 else abs(k2 - 3042.58677467) <= tol * 3042.6
    return {"unittests.py:95": {"got": k1, "want": want1, "ok": bool(abs(k1 - want1) <= tol * want1)},
            "unittests.py:101": {"got": k2, "want": 3042.58677467, "ok": bool(ok2)}}


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[0], [1], [2], [4]: the per-proposal shape of kamphir.py:434-450 (2 DPs per replicate, :290-291)
# ---------------------------------------------------------------------------------------------------
def proposal_entry(name, obs, sims, precision, n_states=0, cpu_reps=4, parity_reps=16, score_reps=2, reps=7,
                   cpu=True, note=None):
    """One ABC proposal through PhyloKernel.score_batch: Newick bytes in, scores out."""
    import numpy as np
    from kamphir_b200 import FlatTree, Forest, PhyloKernel, algorithmic_bytes, get_context
    from oracle import c_oracle
    from oracle import phylok_oracle as po
    pk = PhyloKernel(decayFactor=LAM, gaussFactor=TAU, sigma=SIGMA, normalize="mean", precision=precision,
                     n_states=n_states)
    ctx = get_context()
    raw = [s.encode() for s in sims]
    nreps = len(raw)
    fo = FlatTree.from_newick(obs, 3, "mean", n_states)             # the target tree is prepared once per run (kamphir.py:153-157)
    use_kcoal = True
    for _ in range(2):
        res = pk.score_batch(fo, raw, use_kcoal=use_kcoal)
    lat, kms = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = pk.score_batch(fo, raw, use_kcoal=use_kcoal)          # flatten + pack + H2D + one launch + D2H + kcoal
        lat.append(time.perf_counter() - t0)
        kms.append(ctx.last_stats()["ms"])
    st = ctx.last_stats()
    e2e_s, kernel_ms = median(lat), median(kms)
    # exact work of the proposal's launch (M, R counted by the library from per-tree class histograms)
    flats = FlatTree.many_from_newick_list(raw, 3, "mean", n_states, 0)
    forest = Forest(ctx, [fo] + flats, precision)
    pairs = [(0, s + 1) for s in range(nreps)] + [(s + 1, s + 1) for s in range(nreps)]
    stats = forest.pair_stats(forest, pairs)
    nint = [fo.n_internal] + [f.n_internal for f in flats]
    stats["node_sum"] = sum(nint[a] + nint[b] for a, b in pairs)
    h2d = forest.nbytes()
    forest.close()
    alg = algorithmic_bytes(stats, precision)
    peak, peak_src = hbm_peak()
    achieved = alg / (kernel_ms / 1e3) / 1e9
    # parity: K(obs, sim), K(sim, sim) and the normalised value against the C oracle; the final score (kcoal included)
    # against the restated kamphir.py:249-306 on a few replicates
    tol = 1e-9 if precision == 64 else 1e-5

    def cflat(nwk):
        if n_states:
            return c_oracle.labelled_cflat(po.prep_kamphir(po.read_newick(nwk), "mean", po.state_from_name), n_states)
        return c_oracle.CFlat(po.prep_kamphir(po.read_newick(nwk), "mean"))

    co = cflat(obs)
    ref_denom = c_oracle.kernel(co, co, LAM, TAU, SIGMA)
    errs = [abs(res["ref_denom"] - ref_denom) / ref_denom]
    idx = list(range(0, nreps, max(1, nreps // parity_reps)))[:parity_reps]
    for s in idx:
        cs = cflat(sims[s])
        k, d = c_oracle.kernel(co, cs, LAM, TAU, SIGMA), c_oracle.kernel(cs, cs, LAM, TAU, SIGMA)
        errs += [abs(res["k"][s] - k) / k, abs(res["denom"][s] - d) / d,
                 abs(res["knorm"][s] - k / math.sqrt(ref_denom * d)) / (k / math.sqrt(ref_denom * d))]
    score_err = None
    if n_states == 0 and score_reps:
        mean, scores = po.kamphir_evaluate(sims[:score_reps], obs, decayFactor=LAM, gaussFactor=TAU, sigma=SIGMA,
                                           normalize="mean")
        score_err = max(abs(a - b) / b for a, b in zip(res["score"][:score_reps], scores))
    parity = {"replicates_checked": len(idx), "max_rel_err": max(errs), "score_max_rel_err": score_err, "tol": tol,
              "checker": "C oracle on K(obs,sim), K(sim,sim), K(obs,obs), knorm" +
                         ("; restated kamphir.py:249-306 on the score" if score_err is not None else "") +
                         ("; labelled mode: build's own extension oracle (parity unpinned by the reference)" if n_states else ""),
              "ok": bool(max(errs) <= tol and (score_err is None or score_err <= 10 * tol)
                         and np.all(np.isfinite(res["score"])))}
    entry = {
        "workload": name, "n_tips_obs": fo.n_tips, "nreps": nreps, "dps_per_proposal": 2 * nreps, "n_states": n_states,
        "value": nreps / (kernel_ms / 1e3), "unit": "pair-evals/s",
        "what": "replicates scored per second of DP-kernel time (2 DPs each), forest resident",
        "kernel_ms": kernel_ms, "cells_per_s": stats["cells"] / (kernel_ms / 1e3),
        "node_pairs_per_s": stats["node_pairs"] / (kernel_ms / 1e3),
        "e2e": {"value": nreps / e2e_s, "unit": "pair-evals/s", "ms_per_proposal": 1e3 * e2e_s,
                "h2d_bytes_per_step": int(h2d + 16 * nreps), "d2h_bytes_per_step": int(16 * nreps + 8),
                "what": "PhyloKernel.score_batch(obs, Newick byte strings, use_kcoal=True): parse + prep + pack + H2D + "
                        "one launch + D2H + kcoal + scores, median of %d" % reps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                     "cells_per_launch": stats["cells"], "child_reads_per_launch": stats["child_reads"],
                     "dps_per_launch": stats["dps"],
                     "note": "one launch of %d DPs on %d resident CTAs: %s" %
                             (stats["dps"], st["grid"], "latency-bound, the HBM fraction only says so"
                              if stats["dps"] < 4 * st["grid"] else "several waves")},
        "launch": {k: st[k] for k in ("path", "grid", "threads", "smem_bytes")},
        "parity": parity,
    }
    if note:
        entry["note"] = note
    if cpu:
        from oracle import ref_runner
        m = min(cpu_reps, nreps)
        if ref_runner.available():
            # the reference's kernel() has no tip-state mode: the CPU figure is always the unlabelled kernel on these trees
            rk = ref_runner.make_kernel(LAM, TAU, SIGMA, "mean")
            t_obs = ref_runner.prepare(rk, obs, "mean")
            t_sims = [ref_runner.prepare(rk, s, "mean") for s in sims[:m]]
            t0 = time.perf_counter()
            ks = [(rk.kernel(t_obs, t), rk.kernel(t, t)) for t in t_sims]        # kamphir.py:290-291
            dt = time.perf_counter() - t0
            kind = "reference"
        else:
            f_obs = po.prep_kamphir(po.read_newick(obs), "mean")
            f_sims = [po.prep_kamphir(po.read_newick(s), "mean") for s in sims[:m]]
            t0 = time.perf_counter()
            ks = [(po.kernel_arrays(f_obs, f, LAM, TAU, SIGMA), po.kernel_arrays(f, f, LAM, TAU, SIGMA)) for f in f_sims]
            dt = time.perf_counter() - t0
            kind = "port"
        entry["cpu_baseline"] = {"value": m / dt, "unit": "pair-evals/s", "cores": 1, "kind": kind,
                                 "sample": "%d replicates (2 kernel() calls each, kamphir.py:290-291), %.1f s, prep excluded"
                                           % (m, dt)}
        if n_states == 0:
            cpu_err = max(abs(res["k"][s] - ks[s][0]) / ks[s][0] for s in range(m))
            entry["cpu_baseline"]["gpu_vs_cpu_max_rel_err"] = cpu_err
            parity["ok"] = bool(parity["ok"] and cpu_err <= tol)
    return entry


def gram_probe(n_trees, n_tips, precision, seed, n_states=0, prevalence=None):
    """Device-only steady-state figure for a tree size: Gram of n_trees trees, forest resident, best of 3 launches."""
    import ctypes
    from kamphir_b200 import FlatTree, Forest, _lib, algorithmic_bytes, get_context, synth
    from kamphir_b200._lib import check, pk_params
    nwk = synth.tree_set(n_trees, n_tips, seed, "mixed", n_states, prevalence)
    flats = FlatTree.many_from_newick("\n".join(nwk), 3, "mean", n_states)
    ctx = get_context()
    L = _lib.lib()
    d_out, d_diag = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.pk_device_alloc(ctx._h, 8 * n_trees * n_trees, ctypes.byref(d_out)))
    check(L.pk_device_alloc(ctx._h, 8 * n_trees, ctypes.byref(d_diag)))
    forest = Forest(ctx, flats, precision)
    st = forest.gram_stats(32)
    st["node_sum"] = 2 * sum(f.n_internal for f in flats) * (n_trees + 1) // 2
    p = pk_params(LAM, TAU, SIGMA, precision, 0)
    best = 1e30
    for _ in range(3):
        check(L.pk_forest_gram_part(ctx._h, forest._h, ctypes.byref(p), 1, 32, 0, 1, d_out, n_trees, d_diag))
        best = min(best, ctx.last_stats()["ms"])
    last = ctx.last_stats()
    forest.close()
    check(L.pk_device_free(ctx._h, d_out))
    check(L.pk_device_free(ctx._h, d_diag))
    alg = algorithmic_bytes(st, precision)
    peak, _ = hbm_peak()
    return {"what": "Gram of %d x %d-tip trees, forest resident, best of 3 launches" % (n_trees, n_tips),
            "dps": st["dps"], "kernel_ms": best, "pair_evals_per_s": st["dps"] / (best / 1e3),
            "cells_per_s": st["cells"] / (best / 1e3), "algorithmic_GBps": alg / (best / 1e3) / 1e9,
            "frac_of_measured_hbm_peak": alg / (best / 1e3) / 1e9 / peak,
            "launch": {k: last[k] for k in ("path", "grid", "threads", "smem_bytes")}}


def bench_configs(args):
    from kamphir_b200 import synth
    out = {}
    test_tree = open(os.path.join(REPO, "tests", "golden", "trees", "test.tree")).read()
    obs0 = test_tree.split(";")[0].strip() + ";"
    sims0 = [synth.kingman_newick(100, i) for i in range(100)]                   # seed = 0 * 1000 + tree index
    out["0"] = proposal_entry("configs[0]: tests/test.tree[0] vs 100 synthetic 100-tip coalescent trees", obs0, sims0,
                              args.precision, cpu_reps=100, parity_reps=100, score_reps=4)
    obs1 = synth.kingman_newick(200, 1000 + 999)
    sims1 = [synth.kingman_newick(200, 1000 + i) for i in range(100)]
    out["1"] = proposal_entry("configs[1]: SI-model sizing (settings/example.SI.json), 100 simulated 200-tip trees vs "
                              "observed", obs1, sims1, args.precision, cpu_reps=50, parity_reps=32, score_reps=2)
    # configs[2]: multi-type trees, tip names "<i>_<deme>" with rcolgem's 1-based demes (rcolgem.py:274); DiffRisk has two
    # demes with p = 0.5 (settings/example.DiffRisk.json:20-28), Stages three
    sub = {}
    for S, prev, model in ((2, [0.5, 0.5], "DiffRisk"), (3, [0.2, 0.5, 0.3], "Stages")):
        one = lambda s: re.sub(r"_(\d+)", lambda m: "_%d" % (int(m.group(1)) + 1), s)      # noqa: E731
        obs2 = one(synth.kingman_newick(500, 2000 + 999 + S, states=S, prevalence=prev))
        sims2 = [one(synth.kingman_newick(500, 2000 + 100 * S + i, states=S, prevalence=prev)) for i in range(100)]
        for on in (True, False):
            key = "S%d_matching_%s" % (S, "on" if on else "off")
            sub[key] = proposal_entry("configs[2]: %s-style 500-tip trees, %d tip states, state matching %s, 100 replicates"
                                      % (model, S, "on" if on else "off"), obs2, sims2, args.precision,
                                      n_states=S if on else 0, cpu_reps=8, parity_reps=16, score_reps=0,
                                      cpu=not on,
                                      note=None if not on else "labelled mode is an extension with no reference "
                                           "implementation: the parity verdict is against the build's own extension oracle")
        sub["S%d_gram_probe_matching_on" % S] = gram_probe(256, 500, args.precision, 2000, S, prev)
    out["2"] = sub
    obs4 = synth.kingman_newick(2000, 4000 + 999)
    sims4 = [synth.kingman_newick(2000, 4000 + i) for i in range(256)]
    out["4"] = proposal_entry("configs[4]: PANGEA-scale (pangea_scA.json sizing), 256 simulated 2000-tip trees vs observed, "
                              "C matrix in HBM slabs", obs4, sims4, args.precision, cpu_reps=2, parity_reps=16,
                              score_reps=1, reps=5)
    out["4"]["gram_probe"] = gram_probe(192, 2000, args.precision, 4000)
    return out


def kernel_call_latency(precision):
    """PhyloKernel.kernel(tree object, tree object): the call kamphir.py:290-291 makes, per call, on Bio.Phylo-shaped
    objects (kamphir_b200.synth.Tree).  'cached' = both trees seen before (the observed tree of an MCMC run);
    'fresh_sim' = kernel(obs, sim) + kernel(sim, sim) on a simulated tree object never seen before, per call."""
    from kamphir_b200 import PhyloKernel, synth
    pk = PhyloKernel(decayFactor=LAM, gaussFactor=TAU, sigma=SIGMA, precision=precision)
    out = {}
    for tips in (100, 200, 500):
        obs = synth.newick_to_tree(synth.kingman_newick(tips, 7000 + tips))
        pk.annotate_tree(obs)
        sims = [synth.newick_to_tree(synth.kingman_newick(tips, 7100 + tips + i)) for i in range(24)]
        for t in sims:
            pk.annotate_tree(t)
        pk.kernel(obs, sims[0])
        pk.kernel(sims[0], sims[0])
        t0 = time.perf_counter()
        for _ in range(200):
            pk.kernel(obs, sims[0])
        cached = (time.perf_counter() - t0) / 200
        t0 = time.perf_counter()
        for t in sims[4:]:
            pk.kernel(obs, t)
            pk.kernel(t, t)
        fresh = (time.perf_counter() - t0) / (2 * len(sims[4:]))
        out[str(tips)] = {"cached_us": 1e6 * cached, "fresh_sim_us_per_call": 1e6 * fresh}
    out["what"] = ("per-call latency of PhyloKernel.kernel on tree objects; cached: flattened form kept on the object and "
                   "its single-tree forest on the device; fresh_sim: includes the Python walk of a new 2n-1-node object")
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kamphir_b200 import FlatTree, algorithmic_bytes
    from kamphir_b200.gram import GramRunner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d"
                         % (args.gpus, world, args.gpus))
    # stdout carries the one JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints
    # its version there whenever NCCL_DEBUG is set in the environment) are pointed at stderr, the line goes to a private
    # duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("PK_DEVICE", str(local))
    torch.cuda.set_device(local)
    if world > 1:
        if "PK_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["PK_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    literals = check_literals(args.precision) if rank == 0 else None

    t0 = time.perf_counter()
    newicks = make_trees(args.trees, args.tips)
    t_gen = time.perf_counter() - t0
    text = "\n".join(newicks).encode()      # Newick bytes as a simulator hands them over; the C++ flattener is what is timed
    t0 = time.perf_counter()
    flats = FlatTree.many_from_newick(text, 3, "mean", 0, 0)
    t_flat = time.perf_counter() - t0
    t0 = time.perf_counter()                # again: the first call of a process also pays for the growth of the workers' heaps
    again = FlatTree.many_from_newick(text, 3, "mean", 0, 0)
    t_flat_warm = time.perf_counter() - t0
    del again, text

    runner = GramRunner(flats, LAM, TAU, SIGMA, args.precision, True, args.tile, args.gather, rank, world, local)
    stream = torch.cuda.Stream()            # a real (non-legacy-default) stream shared by torch events and libphylok
    torch.cuda.set_stream(stream)
    runner.ctx.set_stream(stream.cuda_stream)
    forest_bytes = runner.upload()
    stats = runner.my_stats()
    n = args.trees
    total_dps = n * (n + 1) // 2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        runner.launch()
        runner.finish()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    launches0 = runner.ctx.last_stats()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations, outside the per-step event pair
        if world > 1:
            dist.barrier()
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    last = runner.ctx.last_stats()
    launches = last["launches"] - launches0
    t_total = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_total.item()) / args.steps
    value = total_dps / (ms_per_step / 1e3)

    # dominant kernel: the Gram DP launch of this rank, timed by the library's own event pair around that launch
    dp_ms = []
    for _ in range(3):
        flush.zero_()
        runner.launch()
        dp_ms.append(runner.ctx.last_stats()["ms"])
        runner.finish()
    dp_ms_avg = sum(dp_ms) / len(dp_ms)
    alg_bytes = algorithmic_bytes(stats, args.precision)

    # ---- end to end through host buffers: packed trees H2D + all launches + matrix D2H, per step; median of N -------
    barrier()
    e2e_times, parts = [], []
    result = None
    for s in range(max(1, args.e2e_reps)):
        barrier()
        t0 = time.perf_counter()
        h2d = runner.upload()
        t1 = time.perf_counter()
        step()
        t2 = time.perf_counter()
        if rank == 0:
            result = runner.fetch()
        t3 = time.perf_counter()
        barrier()
        t4 = time.perf_counter()
        t = torch.tensor([t4 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)        # the step is over when the slowest rank is done
        e2e_times.append(float(t.item()))
        parts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    e2e_s = median(e2e_times)

    status = 0
    if rank == 0:
        if args.corrupt:
            result[3, 5] *= 1.0 + 1e-6
        # ---- parity gate on a sample (C oracle; the Python port is checked against it in tests/) --------------
        pairs = sample_pairs(n, args.parity_sample, seed=5)
        c_rate, c_ks = c_oracle_rate(newicks, pairs + [(a, a) for a, _ in pairs] + [(b, b) for _, b in pairs])
        m = len(pairs)
        errs = []
        for idx, (a, b) in enumerate(pairs):
            want = c_ks[idx] / math.sqrt(c_ks[m + idx] * c_ks[2 * m + idx])
            errs.append(max(abs(result[a, b] - want), abs(result[b, a] - want)) / want)
        tol = 1e-9 if args.precision == 64 else 1e-5
        parity = {"checked": m, "max_rel_err": max(errs), "tol": tol, "ok": bool(max(errs) <= tol),
                  "symmetric": bool(np.array_equal(result, result.T)),
                  "unit_diagonal": bool(np.allclose(np.diag(result), 1.0, rtol=0, atol=1e-12)),
                  # a normalised kernel matrix: every entry finite, positive and (Cauchy-Schwarz) at most 1
                  "range_ok": bool(np.all(np.isfinite(result)) and result.min() > 0.0 and result.max() <= 1.0 + 1e-12),
                  "reference_literals": literals}
        gate = (parity["ok"] and parity["symmetric"] and parity["unit_diagonal"] and parity["range_ok"]
                and all(v["ok"] for v in literals.values()))
        if not gate:
            print("PARITY FAILURE (exit status 3): %r" % (parity,), file=sys.stderr)
            status = 3

        peak, peak_src = hbm_peak()
        achieved = alg_bytes / (dp_ms_avg / 1e3) / 1e9
        # DRAM traffic of the same kernel from the committed ncu --set full capture, scaled per DP to this launch
        traffic, traffic_src = None, None
        for rnd in ("r02e", "r02d", "r02b", "r02", "r01"):
            try:
                cap_name = "%s_dram_traffic%s.json" % (rnd, "_fp32" if args.precision == 32 else "")
                cap = json.load(open(os.path.join(REPO, "profiles", cap_name)))
                if cap.get("n_tips") == args.tips and cap.get("precision") == args.precision:
                    traffic = cap["dram_bytes_per_dp"] * stats["dps"]
                    traffic_src = ("profiles/%s: (dram__bytes_read.sum + dram__bytes_write.sum) / %d DPs "
                                   "of one captured Gram launch x %d DPs of this launch"
                                   % (cap_name, cap["dps_in_launch"], stats["dps"]))
                    break
            except (OSError, ValueError, KeyError):
                pass
        mid = sorted(range(len(e2e_times)), key=lambda k: e2e_times[k])[len(e2e_times) // 2]
        line = {
            "metric": "normalised-kernel pair-evals/s (Gram matrix, 500-tip trees)", "value": value,
            "unit": "pair-evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "config": workload_config(args, gather=runner.gather,
                                      l2="256 MiB device memset between timed steps (outside the per-step event pair); "
                                         "working set (forest + C slabs + matrix) also exceeds the 126 MB L2",
                                      launch={k: last[k] for k in ("path", "grid", "threads", "smem_bytes")}),
            "pair_evals_per_step": total_dps,
            "node_pairs_per_s": (args.tips - 1) ** 2 * value,
            "cells_per_s_rank0": stats["cells"] / (dp_ms_avg / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "pk_dp_kernel (Gram tile launch of rank 0)", "kernel_ms": dp_ms_avg,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "per_dp": "w*(M+R) + 32*(n1+n2) + 8 (SURVEY.md 8d), M and R counted exactly",
                         "dps_per_launch": stats["dps"], "cells_per_launch": stats["cells"],
                         "child_reads_per_launch": stats["child_reads"]},
            "e2e": {"value": total_dps / e2e_s, "unit": "pair-evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(8 * n * n), "seconds_per_step": e2e_s,
                    "repetitions": len(e2e_times), "statistic": "median (max over ranks per repetition)",
                    "all_seconds": e2e_times,
                    "rank0_breakdown_s": dict(zip(("upload", "launch_finish", "fetch", "closing_barrier"), parts[mid])),
                    "what": "GramRunner.upload() + launch + finish + fetch(): packed host trees -> host N x N matrix"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "host_prep": {"generate_s": t_gen, "flatten_s": t_flat, "flatten_trees_per_s": n / t_flat,
                          "flatten_second_call_s": t_flat_warm, "flatten_second_call_trees_per_s": n / t_flat_warm,
                          "what": "Newick bytes -> kernel-ordered host trees (C++ flattener, all host threads of this rank)"},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_configs:
            cfgs = bench_configs(args)
            line["configs"] = cfgs
            line["configs"]["3"] = {"workload": line["config"]["workload"], "value": value, "unit": "pair-evals/s",
                                    "e2e": {"value": line["e2e"]["value"], "unit": "pair-evals/s"},
                                    "roofline_frac": achieved / peak, "parity": {"ok": bool(gate)},
                                    "note": "the headline: every top-level key of this line"}
            line["kernel_call"] = kernel_call_latency(args.precision)

            def all_ok(d):
                if isinstance(d, dict):
                    if "parity" in d and not d["parity"].get("ok", True):
                        return False
                    return all(all_ok(v) for v in d.values())
                return True
            if not all_ok(cfgs):
                print("PARITY FAILURE in the configs block (exit status 3)", file=sys.stderr)
                status = 3
        if world == 1 and not args.no_cpu_baseline:
            cpairs = sample_pairs(n, args.cpu_sample, seed=9)
            rate, ks, dt = cpu_baseline_single(newicks, cpairs)
            c_rate2, c_ks2 = c_oracle_rate(newicks, cpairs)
            cpu_err = max(abs(a - b) / b for a, b in zip(ks, c_ks2))
            line["cpu_baseline"] = {
                "value": rate, "unit": "pair-evals/s", "cores": 1, "kind": cpu_kind(),
                "sample": "%d random (i<=j) pairs of the %d trees, single process, %.1f s; kernel() only, prep excluded"
                          % (len(cpairs), n, dt),
                "note": cpu_note(),
                "python_port_pair_evals_per_s": port_rate(newicks, cpairs[:64]),
                "c_restatement_pair_evals_per_s": c_rate2, "cpu_vs_c_oracle_max_rel_diff": cpu_err,
                "host_cores_available": os.cpu_count()}
        print(json.dumps(line), file=json_out, flush=True)
    barrier()
    runner.close()
    if world > 1:
        st = torch.tensor([status], dtype=torch.int32, device="cuda")
        dist.broadcast(st, src=0)
        status = int(st.item())
        dist.destroy_process_group()
    if status:
        sys.exit(status)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
