"""world_size-2 gloo test (CPU) of the sharded Gram path's host logic: the tile partition every rank derives from
pk_gram_part_tiles is disjoint and complete, and summing the ranks' blocks (what gather="allreduce" does with NCCL, and
what the P2P direct-write assembles in rank 0's buffer) reproduces the full matrix.  The per-entry values come from the
C oracle here -- the CUDA kernel cannot run on this box; tests/test_gpu_parity.py::test_gram_parts_tile_the_matrix is the
device-side twin."""
import os
import socket

import numpy as np
import pytest

from conftest import REPO  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, tile, result_dir):
    import torch
    import torch.distributed as dist
    from kamphir_b200 import synth
    from kamphir_b200.gram import part_pairs, part_tiles
    from oracle import c_oracle
    from oracle import phylok_oracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nwk = synth.tree_set(n, 14, 700, "mixed")
    flats = [c_oracle.CFlat(po.prep_kamphir(po.read_newick(s))) for s in nwk]
    mine = np.zeros((n, n))
    count = 0
    for ti, tj in part_tiles(n, tile, rank, world):
        for i in range(ti * tile, min(n, (ti + 1) * tile)):
            for j in range(max(i, tj * tile), min(n, (tj + 1) * tile)):
                k = c_oracle.kernel(flats[i], flats[j], 0.2, 2.0)
                assert mine[i, j] == 0.0
                mine[i, j] = mine[j, i] = k
                count += 1
    assert count == part_pairs(n, tile, rank, world)
    total = torch.tensor([count], dtype=torch.int64)
    dist.all_reduce(total)
    full = torch.from_numpy(mine.copy())
    overlap = torch.from_numpy((mine != 0).astype(np.int64))
    dist.all_reduce(full)
    dist.all_reduce(overlap)
    if rank == 0:
        want = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                want[i, j] = want[j, i] = c_oracle.kernel(flats[i], flats[j], 0.2, 2.0)
        np.save(os.path.join(result_dir, "ok.npy"),
                np.array([int(total.item()) == n * (n + 1) // 2, bool(np.array_equal(full.numpy(), want)),
                          int(overlap.max().item()) == 1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,tile", [(13, 4), (9, 1), (16, 8)])
def test_two_ranks_tile_partition_and_sum(tmp_path, n, tile):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), n, tile, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all(), ok


def test_partition_counts_for_bench_shape():
    from kamphir_b200.gram import part_pairs, part_tiles
    n, tile = 4096, 32
    for world in (1, 2, 4, 8):
        pairs = [part_pairs(n, tile, r, world) for r in range(world)]
        assert sum(pairs) == n * (n + 1) // 2
        assert max(pairs) - min(pairs) <= tile * tile * 2      # round-robin over 8256 tiles: near-perfect balance
        tiles = np.concatenate([part_tiles(n, tile, r, world) for r in range(world)])
        assert len(tiles) == 128 * 129 // 2 and len({(int(a), int(b)) for a, b in tiles}) == len(tiles)


def _score_worker(rank, world, port, n, result_dir):
    import torch.distributed as dist
    from kamphir_b200 import synth
    from kamphir_b200.score import shard_indices, sharded_scores
    from oracle import phylok_oracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    obs = synth.kingman_newick(20, 1)
    sims = [synth.kingman_newick(20, 50 + i) for i in range(n)]

    def score_fn(part):      # the oracle's restatement of kamphir.py:249-306 stands in for the CUDA path on this box
        mean, scores = po.kamphir_evaluate(part, obs, decayFactor=0.2, gaussFactor=2.0, normalize="mean")
        return dict(score=np.array(scores), knorm=np.array(scores) * 0.5, ref_denom=1.25)

    res = sharded_scores(score_fn, sims, rank, world)
    assert sorted(sum((shard_indices(n, r, world) for r in range(world)), [])) == list(range(n))
    if rank == 1:            # every rank holds the full result
        want_mean, want = po.kamphir_evaluate(sims, obs, decayFactor=0.2, gaussFactor=2.0, normalize="mean")
        np.save(os.path.join(result_dir, "ok.npy"),
                np.array([bool(np.array_equal(res["score"], np.array(want))), abs(res["mean_score"] - want_mean) < 1e-15,
                          res["ref_denom"] == 1.25, bool(np.array_equal(res["knorm"], np.array(want) * 0.5)),
                          "kcoal" not in res]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8, 1])
def test_two_ranks_sharded_proposal_scores(tmp_path, n):
    """SURVEY 8e, sim-vs-obs row: replicates dealt to the ranks, one all-gather, original order restored (odd counts and
    a rank without work included)."""
    import torch.multiprocessing as mp
    mp.spawn(_score_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all(), ok


def _failing_score_worker(rank, world, port, result_dir):
    import torch.distributed as dist
    from kamphir_b200.score import sharded_scores

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def score_fn(part):      # rank 1's share holds a replicate with K = 0 (math.log(0) at kamphir.py:300)
        if rank == 1:
            raise ValueError("math domain error (kamphir.py:300): K = 0 for replicate 0")
        return dict(score=np.ones(len(part)), knorm=np.ones(len(part)), ref_denom=1.0)

    raised = False
    try:
        sharded_scores(score_fn, ["t%d" % i for i in range(5)], rank, world)
    except ValueError:
        raised = True
    np.save(os.path.join(result_dir, "raised%d.npy" % rank), np.array([raised]))
    dist.barrier()           # reached by both ranks: nobody is left waiting in the all-gather
    dist.destroy_process_group()


def test_two_ranks_raise_together_when_one_share_fails(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_failing_score_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "raised0.npy").all() and np.load(tmp_path / "raised1.npy").all()
