"""Gram-matrix driver over the GPUs of one box: the working, sharded form of ``PhyloKernel.compute_matrix``
(``/root/reference/phyloK2.py:148-156``).

Every (i, j) entry is an independent DP, so the upper triangle is cut into tile x tile blocks that are dealt
to the ranks by cost (node pairs per tile; round-robin for equal-sized trees) (one process per GPU, ``torch.distributed``).  Each rank holds the whole packed forest
(tens of MB) and evaluates its blocks with no data-path collective.  The only thing that crosses NVLink is the
result:

* ``gather="p2p"`` (default for world > 1): rank 0 owns the N x N matrix; the other ranks map it through CUDA IPC and
  their DP kernels store each K (and its mirror) straight into rank 0's HBM -- the result crosses NVLink exactly once,
  written by the kernel that produced it, and there is no gather step at all (a barrier closes the step).
* ``gather="allreduce"``: every rank fills its own zero-initialised matrix and one NCCL all-reduce (sum with zeros is
  exact) leaves the full matrix on every rank.

torch is used for the process group, streams and events only.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, pk_params
from .phylokernel import Forest, get_context

__all__ = ["GramRunner", "part_tiles", "part_pairs"]


class GramRunner(object):
    def __init__(self, flats, decayFactor=0.2, gaussFactor=2.0, sigma=1.0, precision=64, normalised=True, tile=32,
                 gather="p2p", rank=0, world=1, device=None, group=None):
        self.flats = list(flats)
        self.n = len(self.flats)
        self.params = pk_params(float(decayFactor), float(gaussFactor), float(sigma), int(precision), 0)
        self.precision = int(precision)
        self.normalised = bool(normalised)
        self.tile = int(tile)
        self.rank, self.world = int(rank), int(world)
        self.gather = gather if self.world > 1 else "local"
        self.group = group
        self.ctx = get_context(device)
        self.L = _lib.lib()
        self.forest = None
        self._handles = Forest.handle_array(self.flats)
        self.d_out = ctypes.c_void_p()        # where this rank's kernel writes (local or rank 0's mapped buffer)
        self.d_local = ctypes.c_void_p()      # buffer this rank owns (rank 0 always; all ranks for allreduce)
        self.d_diag = ctypes.c_void_p()
        self.h_out = ctypes.c_void_p()        # page-locked host copy of the matrix (fetch)
        self._torch_view = None
        n = self.n
        check(self.L.pk_device_alloc(self.ctx._h, 8 * n, ctypes.byref(self.d_diag)))
        if self.gather in ("local", "allreduce") or self.rank == 0:
            check(self.L.pk_device_alloc(self.ctx._h, 8 * n * n, ctypes.byref(self.d_local)))
            check(self.L.pk_memset_d(self.ctx._h, self.d_local, 0, 8 * n * n))
            self.d_out = self.d_local
        if self.gather == "p2p":
            self._share_rank0_buffer()

    # -- setup ----------------------------------------------------------------------------------------
    def _share_rank0_buffer(self):
        import torch
        import torch.distributed as dist
        handle = torch.zeros(64, dtype=torch.uint8)
        if self.rank == 0:
            raw = (ctypes.c_ubyte * 64)()
            check(self.L.pk_ipc_export(self.ctx._h, self.d_local, raw))
            handle = torch.tensor(list(raw), dtype=torch.uint8)
        dev = torch.device("cuda", torch.cuda.current_device())
        h = handle.to(dev)
        dist.broadcast(h, src=self._global_rank(0), group=self.group)
        if self.rank != 0:
            raw = (ctypes.c_ubyte * 64)(*h.cpu().tolist())
            check(self.L.pk_ipc_open(self.ctx._h, raw, ctypes.byref(self.d_out)))

    def _global_rank(self, r):
        """torch.distributed takes GLOBAL ranks as src/dst, also for a sub-group."""
        if self.group is None:
            return r
        import torch.distributed as dist
        return dist.get_global_rank(self.group, r)

    def upload(self):
        """Host trees -> packed device forest on every rank (the H2D leg of the end-to-end path).  Returns the bytes
        this rank copied from the host.

        With W ranks each one packs and uploads 1/W of the trees and the ranks broadcast their byte ranges to each other
        device-to-device (NCCL over NVLink): the host packs the forest once per box instead of once per GPU."""
        import os
        import time
        trace = bool(os.environ.get("PK_TRACE"))
        t0 = time.perf_counter()
        if self.forest is not None:
            self.forest.close()
        t1 = time.perf_counter()
        if self.world == 1:
            self.forest = Forest(self.ctx, self.flats, self.precision, keep=False, handles=self._handles)
            if trace:
                print("GramRunner.upload: close %.0f us, Forest() %.0f us" % (1e6 * (t1 - t0), 1e6 * (time.perf_counter() - t1)),
                      file=__import__("sys").stderr)
            return self.forest.nbytes()
        import torch
        import torch.distributed as dist
        cuts = [self.n * r // self.world for r in range(self.world + 1)]
        self.forest = Forest(self.ctx, self.flats, self.precision, part=(cuts[self.rank], cuts[self.rank + 1]), keep=False,
                             handles=self._handles)
        t2 = time.perf_counter()
        self.ctx.sync()
        t3 = time.perf_counter()
        mine = 0
        for r in range(self.world):
            ptr, nbytes = self.forest.device_range(cuts[r], cuts[r + 1])
            if r == self.rank:
                mine = nbytes
            if nbytes:
                dist.broadcast(_as_tensor(ptr, nbytes, "|u1"), src=self._global_rank(r), group=self.group)
        t4 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        if trace and self.rank == 0:
            print("GramRunner.upload: close %.0f us, Forest(part) %.0f us, sync %.0f us, %d broadcasts queued %.0f us, "
                  "waited %.0f us" % (1e6 * (t1 - t0), 1e6 * (t2 - t1), 1e6 * (t3 - t2), self.world, 1e6 * (t4 - t3),
                                      1e6 * (time.perf_counter() - t4)), file=__import__("sys").stderr)
        return mine + 32 * self.n           # + the meta records every rank uploads

    # -- one step ---------------------------------------------------------------------------------------
    def launch(self):
        """Asynchronously evaluate this rank's blocks (plus the N diagonal DPs every rank needs)."""
        if self.gather == "allreduce":
            self.zero()                          # the all-reduce sums the ranks' disjoint blocks with zeros elsewhere
        check(self.L.pk_forest_gram_part(self.ctx._h, self.forest._h, ctypes.byref(self.params),
                                         1 if self.normalised else 0, self.tile, self.rank, self.world,
                                         self.d_out, self.n, self.d_diag))

    def finish(self):
        """Close the step: after this, rank 0 (p2p) / every rank (allreduce) holds the full matrix in HBM."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        if self.gather == "allreduce":
            if self._torch_view is None:
                self._torch_view = _as_tensor(self.d_local.value, self.n * self.n)
            # the Gram kernels run on the library's stream, NCCL on torch's current stream, and fetch() copies on the
            # library's stream again: order the three explicitly (they are the same stream only if the caller bound them)
            self.ctx.sync()
            dist.all_reduce(self._torch_view, group=self.group)
            torch.cuda.current_stream().synchronize()
        else:
            self.ctx.sync()                      # peer stores issued by this rank's kernel are complete
            dist.barrier(group=self.group)

    def zero(self):
        if self.d_local.value:
            check(self.L.pk_memset_d(self.ctx._h, self.d_local, 0, 8 * self.n * self.n))

    def fetch(self, copy=False):
        """Device -> host copy of the full matrix (rank 0 for p2p; any rank otherwise) into a page-locked buffer that
        this runner owns and reuses: the returned array is a view of it (valid until the next fetch / close) unless
        copy=True."""
        nbytes = 8 * self.n * self.n
        if not self.h_out.value:
            check(self.L.pk_host_alloc(self.ctx._h, nbytes, ctypes.byref(self.h_out)))
        check(self.L.pk_memcpy_d2h(self.ctx._h, self.h_out, self.d_local, nbytes))
        if self.precision == 32 and self.ctx.nonfinite(reset=True):
            # every rank counts on its own device; rank 0 sees its own tiles here, the matrix scan below sees the rest
            raise OverflowError("fp32 cells overflowed in the Gram kernels (decayFactor too large for these trees in "
                                "fp32): use precision=64")
        view = np.ctypeslib.as_array(ctypes.cast(self.h_out, ctypes.POINTER(ctypes.c_double)), shape=(self.n, self.n))
        return view.copy() if copy else view

    def my_stats(self):
        st = self.forest.gram_stats(self.tile, self.rank, self.world)
        st["node_sum"] = sum(self._node_sum_terms())
        return st

    def _node_sum_terms(self):
        # sum over this rank's DPs of (n1 + n2): exact, from the tile assignment
        n, t = self.n, self.tile
        nint = np.array([f.n_internal for f in self.flats], dtype=np.int64)
        for ti, tj in self.my_tiles():
            a = nint[ti * t:min(n, (ti + 1) * t)]
            b = nint[tj * t:min(n, (tj + 1) * t)]
            if ti == tj:
                m = len(a)
                idx = np.triu_indices(m)
                yield int((a[idx[0]] + a[idx[1]]).sum())
            else:
                yield int(a.sum() * len(b) + b.sum() * len(a))

    def my_tiles(self):
        """(row block, column block) of the tiles this rank evaluates for the uploaded forest (cost-weighted deal)."""
        cnt = ctypes.c_int64(0)
        check(self.L.pk_forest_gram_part_tiles(self.forest._h, self.tile, self.rank, self.world, None, ctypes.byref(cnt)))
        out = np.zeros((cnt.value, 2), np.int32)
        if cnt.value:
            check(self.L.pk_forest_gram_part_tiles(self.forest._h, self.tile, self.rank, self.world, out.ctypes.data,
                                                   ctypes.byref(cnt)))
        return out

    def close(self):
        if self.forest is not None:
            self.forest.close()
            self.forest = None
        if self.gather == "p2p" and self.rank != 0 and self.d_out.value:
            check(self.L.pk_ipc_close(self.ctx._h, self.d_out))
            self.d_out = ctypes.c_void_p()
        if self.h_out.value:
            check(self.L.pk_host_free(self.ctx._h, self.h_out))
            self.h_out = ctypes.c_void_p()
        for buf in (self.d_local, self.d_diag):
            if buf.value:
                check(self.L.pk_device_free(self.ctx._h, buf))
        self.d_local = ctypes.c_void_p()
        self.d_diag = ctypes.c_void_p()


def part_tiles(n, tile, part, nparts):
    """(row-block, column-block) coordinates of the Gram tiles that `part` of `nparts` evaluates (host only)."""
    L = _lib.lib()
    cnt = ctypes.c_int64(0)
    check(L.pk_gram_part_tiles(int(n), int(tile), int(part), int(nparts), None, ctypes.byref(cnt)))
    out = np.zeros((cnt.value, 2), np.int32)
    if cnt.value:
        check(L.pk_gram_part_tiles(int(n), int(tile), int(part), int(nparts), out.ctypes.data, ctypes.byref(cnt)))
    return out


def part_pairs(n, tile, part, nparts):
    """DPs of that part (excluding the diagonal pass every part repeats)."""
    cnt = ctypes.c_int64(0)
    check(_lib.lib().pk_gram_part_pairs(int(n), int(tile), int(part), int(nparts), ctypes.byref(cnt)))
    return cnt.value


def _as_tensor(ptr, numel, typestr="<f8"):
    """Zero-copy torch view of a device buffer owned by libphylok (for the NCCL collectives only)."""
    import torch

    class _Holder(object):
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=torch.device("cuda", torch.cuda.current_device()))
