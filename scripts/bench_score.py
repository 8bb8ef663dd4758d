"""Sim-vs-obs batch (BASELINE.json configs[0], [1], [4]): one ABC proposal = nreps simulated trees scored against the
observed tree (kamphir.py:434-450).  Reports the latency of the whole host call (Newick strings in, scores out) and its
pieces, next to the oracle's kamphir_evaluate on one CPU core for a subsample."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, PhyloKernel, get_context, synth  # noqa: E402
from oracle import phylok_oracle as po  # noqa: E402


def run(name, n_tips, nreps, obs_newick=None, cpu_reps=4, precision=64):
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, normalize="mean", precision=precision)
    obs = obs_newick or synth.kingman_newick(n_tips, 1000 * 2)
    sims = [synth.kingman_newick(n_tips, 2000 + i) for i in range(nreps)]
    sims = [x.encode() for x in sims]      # what a Python-2 simulator bridge hands over: byte strings
    # warm-up (context creation, first launch)
    pk.score_batch(obs, sims[:2])
    lat, flat_t = [], []
    fo = FlatTree.from_newick(obs)             # the target tree is prepared once per run (kamphir.py:153-157)
    for _ in range(7):
        t0 = time.perf_counter()
        res = pk.score_batch(fo, sims, use_kcoal=True)       # Newick byte strings in, scores out: one library call
        lat.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        fs = FlatTree.many_from_newick_list(sims, 3, "mean", 0, 0)      # the flattening part alone, for the breakdown
        flat_t.append(time.perf_counter() - t0)
        del fs
    st = get_context().last_stats()
    # CPU: the oracle's restatement of kamphir.py:249-306 on a subsample of the replicates
    t0 = time.perf_counter()
    mean_cpu, scores_cpu = po.kamphir_evaluate([x.decode() for x in sims[:cpu_reps]], obs, decayFactor=0.2, gaussFactor=2.0, normalize="mean")
    cpu_s = (time.perf_counter() - t0) / cpu_reps
    err = float(np.max(np.abs(res["score"][:cpu_reps] - np.array(scores_cpu)) / np.array(scores_cpu)))
    out = dict(config=name, n_tips=n_tips, nreps=nreps, precision=precision,
               proposal_latency_ms=1e3 * min(lat), flatten_ms=1e3 * min(flat_t), kernel_ms=st["ms"],
               proposals_per_s=1.0 / min(lat), pair_evals_per_s=nreps / min(lat),
               cpu_port_s_per_replicate=cpu_s, speedup_vs_cpu_1core=cpu_s * nreps / min(lat), max_rel_err=err,
               launch=dict(path=st["path"], grid=st["grid"], threads=st["threads"]))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg0", "cfg1", "cfg4"]
    if "cfg0" in which:
        obs = open("tests/golden/trees/test.tree").read().split(";")[0] + ";"
        run("configs[0]: tests/test.tree[0] vs 100 Kingman 100-tip trees", 100, 100, obs_newick=obs)
    if "cfg1" in which:
        run("configs[1]: SI-model sizing, 100 simulated 200-tip trees vs observed", 200, 100)
    if "cfg4" in which:
        run("configs[4]: PANGEA-scale, 256 simulated 2000-tip trees vs observed", 2000, 256, cpu_reps=1)
