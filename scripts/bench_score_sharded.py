"""configs[4] (256 simulated 2000-tip trees vs observed) scored over the ranks of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_score_sharded.py
Every rank flattens and scores every N-th replicate, one all-gather returns the scores (kamphir_b200/score.py)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from kamphir_b200 import FlatTree, PhyloKernel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    tips = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nreps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0, normalize="mean", device=local)
    obs = FlatTree.from_newick(synth.kingman_newick(tips, 2000))
    sims = [synth.kingman_newick(tips, 2000 + i).encode() for i in range(nreps)]
    ref = pk.score_batch(obs, sims[:2])["ref_denom"]
    lat = []
    for _ in range(6):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pk.score_batch_sharded(obs, sims, rank, world, ref_denom=ref, use_kcoal=True)
        lat.append(time.perf_counter() - t0)
    t = torch.tensor([min(lat)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        one = pk.score_batch(obs, sims, ref_denom=ref, use_kcoal=True)
        err = float(np.max(np.abs(res["score"] - one["score"]) / one["score"]))
        print(json.dumps(dict(n_tips=tips, nreps=nreps, n_gpus=world, proposal_latency_ms=1e3 * float(t.item()),
                              pair_evals_per_s=nreps / float(t.item()), max_rel_diff_vs_one_gpu=err,
                              mean_score=res["mean_score"])), file=sys.stderr, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
