"""Sim-vs-obs scoring of one ABC proposal over the GPUs of one box (second row).

The replicates of a proposal are independent (``kamphir.py:434-450`` loops over them), so rank r of W scores the
replicates r, r+W, r+2W, ... -- its own host threads flatten them, its own GPU runs their DPs -- and the only exchange
is one all-gather of the per-replicate numbers (a few KB).  That pays for PANGEA-size trees (2000 tips: flattening and
packing 256 replicates is ~10 ms of host work per proposal); for 100-200-tip trees one GPU finishes a proposal in about
a millisecond and the ranks are better used as independent MCMC chains ("replicas only").

torch.distributed is the plumbing (NCCL with CUDA tensors, or gloo with CPU tensors in the CPU tests).
"""
import numpy as np

__all__ = ["shard_indices", "sharded_scores"]

KEYS = ("k", "denom", "knorm", "kcoal", "score")


def shard_indices(n, rank, world):
    """Replicates of rank `rank`: strided, so every rank gets the same count (+-1) whatever the order of the list."""
    return list(range(int(rank), int(n), int(world)))


def sharded_scores(score_fn, sims, rank, world, group=None, device=None):
    """score_fn(sub_list) -> dict like PhyloKernel.score_batch (arrays over the sub-list, plus 'ref_denom').

    Returns the same dict over the WHOLE list, identical on every rank: arrays in the original order, 'mean_score' the
    mean of 'score' (or of 'knorm' when kcoal was not requested), 'ref_denom' from rank 0.
    """
    sims = list(sims)
    n = len(sims)
    if n == 0:
        raise ValueError("no simulated trees (kamphir.py:450 divides by nreps)")
    mine = shard_indices(n, rank, world)
    err = None
    try:
        res = score_fn([sims[i] for i in mine]) if mine else {}
    except ValueError as e:      # e.g. K = 0 for one replicate (math.log(0) at kamphir.py:300): every rank must still
        if world == 1:           # reach the all-gather, then all of them raise
            raise
        res, err = {}, str(e)
    if world == 1:
        return res
    import torch
    import torch.distributed as dist
    per = (n + world - 1) // world                       # slots per rank (the last ranks may leave one unused)
    buf = np.full((len(KEYS) + 2, per), np.nan)
    buf[len(KEYS) + 1, 0] = 0.0 if err is None else 1.0
    for r, key in enumerate(KEYS):
        if key in res:
            buf[r, :len(mine)] = res[key]
    buf[len(KEYS), 0] = res.get("ref_denom", np.nan)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    full = {}
    stacked = np.stack([o.cpu().numpy() for o in out])   # [world, keys + 2, per]
    failed = [w for w in range(world) if stacked[w, len(KEYS) + 1, 0] != 0.0]
    if failed:
        raise ValueError(err if err is not None else "rank %d: a replicate's score is undefined (K = 0)" % failed[0])
    for r, key in enumerate(KEYS):
        col = np.full(n, np.nan)
        for w in range(world):
            idx = shard_indices(n, w, world)
            col[idx] = stacked[w, r, :len(idx)]
        if not np.all(np.isnan(col)):
            full[key] = col
    full["ref_denom"] = float(stacked[0, len(KEYS), 0])
    top = full["score"] if "score" in full else full["knorm"]
    full["mean_score"] = float(top.sum() / n)            # kamphir.py:450
    return full
