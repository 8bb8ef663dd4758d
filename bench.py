#!/usr/bin/env python
"""Benchmark of the phyloK2 tree-kernel hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (oracle port), rank 0 only

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): the all-pairs normalised Gram matrix
of 4096 synthetic 500-tip trees (2048 Kingman + 2048 Yule), lambda=0.2, tau=2.0, sigma=1, kamphir-style prep.  One
"step" = one full Gram matrix = N(N+1)/2 tree-pair DPs, sharded over the ranks by tiles (strong scaling: the matrix is
fixed, more GPUs share it).  `value` = pair evaluations per second with the packed forest already resident in HBM;
`e2e` = the same through host buffers (packed-tree upload H2D, result matrix D2H inside the timed region).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

LAM, TAU, SIGMA = 0.2, 2.0, 1.0
BASE_SEED = 3000            # seed = config index * 1000 + tree index (config index 3)


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
    ap.add_argument("--cpu-sample", type=int, default=512, help="DPs in the single-process cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-sample", type=int, default=64)
    return ap.parse_args()


def make_trees(n_trees, n_tips):
    from kamphir_b200 import synth
    return synth.tree_set(n_trees, n_tips, BASE_SEED, "mixed")


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
# CPU baseline: the oracle port of phyloK2.kernel (pure Python, like the reference) and its C restatement
# ---------------------------------------------------------------------------------------------------
_CPU_FLATS = None


def _cpu_init(newicks):
    global _CPU_FLATS
    from oracle import phylok_oracle as po
    _CPU_FLATS = [po.prep_kamphir(po.read_newick(s), "mean") for s in newicks]


def _cpu_dp(pair):
    from oracle import phylok_oracle as po
    a, b = pair
    t0 = time.perf_counter()
    k = po.kernel_arrays(_CPU_FLATS[a], _CPU_FLATS[b], LAM, TAU, SIGMA)
    return k, time.perf_counter() - t0


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
    """--impl reference: the reference's CPU path on the host cores (pure-Python oracle port of phyloK2.py:158-209,
    the same loop nest and arithmetic as the reference, which cannot travel to the GPU box), all host threads,
    each step a bounded sample of the Gram workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(32 * cores, 256)      # ~1 s of work per step on every core: the pool's dispatch cost is negligible
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
        "config": workload_config(args, sample=sample),
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "pure-Python port of phyloK2.PhyloKernel.kernel under CPython %s; the reference needs "
                                 "Python 2.7 + Biopython, absent from this image" % sys.version.split()[0]},
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
    torch.cuda.set_device(local)
    if world > 1:
        if "PK_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["PK_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t0 = time.perf_counter()
    newicks = make_trees(args.trees, args.tips)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    flats = FlatTree.many_from_newick("\n".join(newicks), 3, "mean", 0, 0)
    t_flat = time.perf_counter() - t0

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
    kernel_ms = []
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations, outside the per-step event pair
        if world > 1:
            dist.barrier()
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
        kernel_ms.append(None)
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

    # ---- end to end through host buffers: packed trees H2D + all launches + matrix D2H, per step -------------
    barrier()
    e2e_times = []
    result = None
    for s in range(2):
        barrier()
        t0 = time.perf_counter()
        h2d = runner.upload()
        step()
        if rank == 0:
            result = runner.fetch()
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = min(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        # ---- parity gate on a sample (C oracle; the Python port is checked against it in tests/) --------------
        pairs = sample_pairs(n, args.parity_sample, seed=5)
        c_rate, c_ks = c_oracle_rate(newicks, pairs + [(a, a) for a, _ in pairs] + [(b, b) for _, b in pairs])
        m = len(pairs)
        errs = []
        for idx, (a, b) in enumerate(pairs):
            want = c_ks[idx] / math.sqrt(c_ks[m + idx] * c_ks[2 * m + idx])
            errs.append(abs(result[a, b] - want) / want)
        tol = 1e-9 if args.precision == 64 else 1e-5
        parity = {"checked": m, "max_rel_err": max(errs), "tol": tol, "ok": bool(max(errs) <= tol),
                  "symmetric": bool(np.array_equal(result, result.T)),
                  "unit_diagonal": bool(np.allclose(np.diag(result), 1.0, rtol=0, atol=1e-12))}
        if not (parity["ok"] and parity["symmetric"] and parity["unit_diagonal"]):
            print("PARITY FAILURE: %r" % (parity,), file=sys.stderr)

        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / (dp_ms_avg / 1e3) / 1e9
        # DRAM traffic of the same kernel from the committed ncu --set full capture, scaled per DP to this launch
        traffic, traffic_src = None, None
        try:
            cap_name = "r01_dram_traffic%s.json" % ("_fp32" if args.precision == 32 else "")
            cap = json.load(open(os.path.join(REPO, "profiles", cap_name)))
            if cap.get("n_tips") == args.tips and cap.get("precision") == args.precision:
                traffic = cap["dram_bytes_per_dp"] * stats["dps"]
                traffic_src = ("profiles/%s: (dram__bytes_read.sum + dram__bytes_write.sum) / %d DPs "
                               "of one captured Gram launch x %d DPs of this launch"
                               % (cap_name, cap["dps_in_launch"], stats["dps"]))
        except (OSError, ValueError, KeyError):
            pass
        node_pairs_total = None
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
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "kernel": "pk_dp_kernel (Gram tile launch of rank 0)", "kernel_ms": dp_ms_avg,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "per_dp": "w*(M+R) + 32*(n1+n2) + 8 (SURVEY.md 8d), M and R counted exactly",
                         "dps_per_launch": stats["dps"], "cells_per_launch": stats["cells"],
                         "child_reads_per_launch": stats["child_reads"]},
            "e2e": {"value": total_dps / e2e_s, "unit": "pair-evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(8 * n * n), "seconds_per_step": e2e_s,
                    "what": "GramRunner.upload() + launch + finish + fetch(): packed host trees -> host N x N matrix"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "host_prep": {"generate_s": t_gen, "flatten_s": t_flat, "flatten_trees_per_s": n / t_flat},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            cpairs = sample_pairs(n, args.cpu_sample, seed=9)
            rate, ks, dt = cpu_baseline_single(newicks, cpairs)
            c_rate2, c_ks2 = c_oracle_rate(newicks, cpairs)
            port_err = max(abs(a - b) / b for a, b in zip(ks, c_ks2))
            line["cpu_baseline"] = {
                "value": rate, "unit": "pair-evals/s", "cores": 1, "kind": "port",
                "sample": "%d random (i<=j) pairs of the %d trees, single process, %.1f s; kernel() only, prep excluded"
                          % (len(cpairs), n, dt),
                "note": "pure-Python port of phyloK2.PhyloKernel.kernel (same loop nest/arithmetic) under CPython %s; "
                        "Python 2.7 + Biopython are absent from the image" % sys.version.split()[0],
                "c_restatement_pair_evals_per_s": c_rate2, "port_vs_c_max_rel_diff": port_err,
                "host_cores_available": os.cpu_count()}
        print(json.dumps(line), file=json_out, flush=True)
    barrier()
    runner.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
