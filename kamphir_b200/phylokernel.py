"""`PhyloKernel`: host-side mirror of the reference class (``/root/reference/phyloK2.py:9-236``) over libphylok.so.

Same constructor keywords, attributes and methods as the reference, so ``kamphir.py`` (``class Kamphir(PhyloKernel)``,
``kamphir.py:34,54``; calls at ``:155-157`` and ``:281-291``) keeps working through a one-line
``phyloK2.py`` stub.  The arithmetic runs on the GPU through the C ABI in ``include/phylok.h``;
nothing here computes a kernel value on the CPU and nothing imports ``oracle/``.

Additive, batched entry points (one launch for many tree pairs):
    flatten / flatten_many    Newick text or Bio.Phylo-like tree -> FlatTree (C++ flattener)
    kernel_batch              K for a list of pairs / all pairs of two lists
    score_batch               kamphir.py:288-306 for one target and nreps simulated trees, incl. kcoal
    gram                      all-pairs matrix (raw or normalised)
"""
import ctypes
import math
import os

import numpy as np

from . import _lib
from ._lib import PK_NORM, PK_PREP_KAMPHIR, PK_PREP_LADDERIZE, PK_PREP_ZERO_ROOT, check, pk_params, pk_tree_info

__all__ = ["PhyloKernel", "FlatTree", "Context", "get_context"]


# ---------------------------------------------------------------------------------------------------
# handles
# ---------------------------------------------------------------------------------------------------
class FlatTree(object):
    """Owner of one ``pk_tree*``: a tree flattened to post-order SoA arrays (annotate_tree, phyloK2.py:132-146)."""

    _INFO = ("n_internal", "n_tips", "n_levels", "n_classes", "n_states", "height", "scale")

    def __init__(self, handle, recipe):
        self._h = ctypes.c_void_p(handle)
        self._recipe = recipe
        self._dev = None       # {(pid, device, precision): single-tree Forest}, see device_forest()

    def __getattr__(self, name):
        # tree facts are fetched from the library on first use (keeps bulk flattening cheap on the Python side)
        if name in FlatTree._INFO:
            info = pk_tree_info()
            check(_lib.lib().pk_tree_get_info(self._h, ctypes.byref(info)))
            for k in FlatTree._INFO:
                self.__dict__[k] = getattr(info, k)
            return self.__dict__[name]
        raise AttributeError(name)

    # -- constructors --------------------------------------------------------------------------------
    @classmethod
    def from_newick(cls, text, prep=PK_PREP_KAMPHIR, normalize="mean", n_states=0):
        if isinstance(text, str):
            text = text.encode("utf-8")
        out = ctypes.c_void_p()
        check(_lib.lib().pk_tree_from_newick(text, len(text), int(prep), PK_NORM[normalize], int(n_states),
                                             ctypes.byref(out)))
        return cls(out.value, ("newick", text, int(prep), normalize, int(n_states)))

    @classmethod
    def many_from_newick(cls, text, prep=PK_PREP_KAMPHIR, normalize="mean", n_states=0, nthreads=0):
        """All trees of a multi-tree Newick text (``Phylo.parse``, kamphir.py:112), flattened by worker threads."""
        if isinstance(text, str):
            text = text.encode("utf-8")
        arr = ctypes.POINTER(ctypes.c_void_p)()
        n = ctypes.c_int32()
        check(_lib.lib().pk_trees_from_newick(text, len(text), int(prep), PK_NORM[normalize], int(n_states),
                                              int(nthreads), ctypes.byref(arr), ctypes.byref(n)))
        try:
            return [cls(arr[i], ("handle",)) for i in range(n.value)]
        finally:
            _lib.lib().pk_tree_array_free(arr)

    @classmethod
    def many_from_newick_list(cls, texts, prep=PK_PREP_KAMPHIR, normalize="mean", n_states=0, nthreads=0):
        """One FlatTree per Newick string of ``texts`` (str or bytes), flattened in one library call by worker threads."""
        raw = [t.encode("utf-8") if isinstance(t, str) else t for t in texts]
        n = len(raw)
        ptrs = (ctypes.c_char_p * n)(*raw)
        lens = (ctypes.c_size_t * n)(*[len(b) for b in raw])
        arr = ctypes.POINTER(ctypes.c_void_p)()
        cnt = ctypes.c_int32()
        check(_lib.lib().pk_trees_from_newick_list(ptrs, lens, n, int(prep), PK_NORM[normalize], int(n_states),
                                                   int(nthreads), ctypes.byref(arr), ctypes.byref(cnt)))
        try:
            return [cls(arr[i], ("newick", raw[i], int(prep), normalize, int(n_states))) for i in range(cnt.value)]
        finally:
            _lib.lib().pk_tree_array_free(arr)

    @classmethod
    def from_arrays(cls, cidx, bl, classes=None, n_classes=3):
        cidx = np.ascontiguousarray(cidx, dtype=np.int32).reshape(-1)
        bl = np.ascontiguousarray(bl, dtype=np.float64).reshape(-1)
        if cidx.size != bl.size or cidx.size % 2:
            raise ValueError("cidx and bl must both have 2 entries per internal node")
        c = None
        if classes is not None:
            c = np.ascontiguousarray(classes, dtype=np.uint8)
        out = ctypes.c_void_p()
        check(_lib.lib().pk_tree_from_arrays(cidx.size // 2, cidx.ctypes.data, bl.ctypes.data,
                                             c.ctypes.data if c is not None else None, int(n_classes),
                                             ctypes.byref(out)))
        return cls(out.value, ("arrays", cidx, bl, c, int(n_classes)))

    # -- queries --------------------------------------------------------------------------------------
    def export(self):
        """(prod[n], cprod[n,2], cidx[n,2], bl[n,2]) in post-order, as annotate_tree defines them."""
        n = self.n_internal
        prod = np.empty(n, np.uint8)
        cprod = np.empty((n, 2), np.uint8)
        cidx = np.empty((n, 2), np.int32)
        bl = np.empty((n, 2), np.float64)
        check(_lib.lib().pk_tree_export(self._h, prod.ctypes.data, cprod.ctypes.data, cidx.ctypes.data, bl.ctypes.data))
        return prod, cprod, cidx, bl

    def node_heights(self):
        h = np.empty(self.n_internal, np.float64)
        check(_lib.lib().pk_tree_node_heights(self._h, h.ctypes.data))
        return h

    def pair_counts(self, other):
        c = (ctypes.c_int64 * 3)()
        check(_lib.lib().pk_pair_counts(self._h, other._h, c))
        return int(c[0]), int(c[1]), int(c[2])

    def kcoal(self, target):
        out = ctypes.c_double()
        check(_lib.lib().pk_kcoal(self._h, target._h, ctypes.byref(out)))
        return out.value

    def device_forest(self, ctx, precision=64):
        """This tree alone as a device-resident forest, uploaded on first use and kept until the tree is dropped:
        ``PhyloKernel.kernel(t1, t2)`` on trees it has seen before is then one launch and one copy of 8 bytes (the
        observed tree of an MCMC run is passed to ``kernel`` 2 * nreps times per step, kamphir.py:290)."""
        key = (os.getpid(), ctx.device, int(precision))
        if self._dev is None:
            self._dev = {}
        f = self._dev.get(key)
        if f is None or f._h is None:
            f = Forest(ctx, [self], precision, keep=False)
            self._dev = {k: v for k, v in self._dev.items() if k[0] == key[0]}     # a forked child drops the parent's
            self._dev[key] = f
        return f

    # -- lifetime / pickling (Kamphir dill-pickles itself into Pool workers, kamphir.py:27-32) ------------
    def __del__(self):
        dev, self._dev = getattr(self, "_dev", None), None
        if dev:
            for f in dev.values():
                try:
                    f.close()
                except Exception:
                    pass
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value:
            try:
                _lib.lib().pk_tree_free(h)
            except Exception:
                pass

    def __reduce__(self):
        r = self._recipe
        if r[0] == "newick":
            return (_rebuild_newick, r[1:])
        prod, cprod, cidx, bl = self.export()
        # what the arrays do not carry (a tree parsed by the library from a multi-tree text keeps no text of its own):
        # node heights for kcoal, un-normalised height, normalisation divisor
        try:
            extra = (self.node_heights(), float(self.height), float(self.scale))
        except ValueError:
            extra = None
        if self.n_states != 0:
            return (_rebuild_arrays, (cidx, bl, prod - 1, self.n_classes, extra))
        return (_rebuild_arrays, (cidx, bl, None, 3, extra))


def _rebuild_newick(text, prep, normalize, n_states):
    return FlatTree.from_newick(text, prep, normalize, n_states)


def _rebuild_arrays(cidx, bl, classes, n_classes, extra=None):
    t = FlatTree.from_arrays(cidx, bl, classes, n_classes)
    if extra is not None:
        h = np.ascontiguousarray(extra[0], dtype=np.float64)
        check(_lib.lib().pk_tree_set_heights(t._h, h.ctypes.data, h.size, extra[1], extra[2]))
    return t


class Context(object):
    """One ``pk_ctx*`` (GPU + stream + scratch).  Created lazily, per process (never before a fork, kamphir.py:638)."""

    def __init__(self, device=0):
        self.device = int(device)
        self.pid = os.getpid()
        out = ctypes.c_void_p()
        check(_lib.lib().pk_ctx_create(self.device, ctypes.byref(out)))
        self._h = out
        threads = int(os.environ.get("PK_THREADS", "0"))
        ctas = int(os.environ.get("PK_CTAS", "0"))
        force = int(os.environ.get("PK_FORCE_HBM", "0"))
        if threads or ctas or force:
            self.configure(threads, ctas, force)
        if os.environ.get("PK_UNSTAGED"):
            self.set_option("unstaged", int(os.environ["PK_UNSTAGED"]))

    def configure(self, threads=0, ctas_per_sm=0, force_hbm=0):
        check(_lib.lib().pk_ctx_configure(self._h, int(threads), int(ctas_per_sm), int(force_hbm)))

    def set_option(self, name, value):
        check(_lib.lib().pk_ctx_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream):
        check(_lib.lib().pk_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream) if cuda_stream else None))

    def sync(self):
        check(_lib.lib().pk_ctx_sync(self._h))

    def last_stats(self):
        ms, launches, path = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int32()
        cfg = (ctypes.c_int32 * 3)()
        check(_lib.lib().pk_ctx_last_stats(self._h, ctypes.byref(ms), ctypes.byref(launches), ctypes.byref(path), cfg))
        return dict(ms=ms.value, launches=launches.value,
                    path={0: "smem", 1: "hbm", 2: "hbm-unstaged", 3: "hbm-colglobal"}.get(path.value, "hbm"),
                    grid=cfg[0], threads=cfg[1], smem_bytes=cfg[2])

    def nonfinite(self, reset=True):
        """Results that came out inf / NaN since the last reset (fp32 cell overflow); waits for the stream."""
        c = ctypes.c_int64()
        check(_lib.lib().pk_ctx_nonfinite(self._h, ctypes.byref(c), 1 if reset else 0))
        return c.value

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value and self.pid == os.getpid():
            _lib.lib().pk_ctx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Forest(object):
    """Device-resident packed SoA of a set of FlatTrees (``pk_forest*``)."""

    def __init__(self, ctx, trees, precision=64, part=None, keep=True, handles=None):
        """part=(lo, hi): pack and upload only those trees (the rest of the device forest is filled by the other ranks,
        see GramRunner.upload); the layout and the statistics always cover the whole set.  keep=False: do not hold on
        to the FlatTrees (the device copy does not need them once it is made).  handles: the ctypes array of the trees'
        handles when the caller uploads the same list again and again (building it costs 0.6 ms for 4096 trees)."""
        self.ctx = ctx
        trees = list(trees) if handles is None or keep else trees
        self.trees = trees if keep else None
        self.n = len(trees)
        self.precision = int(precision)
        arr = handles if handles is not None else Forest.handle_array(trees)
        out = ctypes.c_void_p()
        if part is None:
            check(_lib.lib().pk_forest_upload(ctx._h, arr, self.n, self.precision, ctypes.byref(out)))
        else:
            check(_lib.lib().pk_forest_upload_part(ctx._h, arr, self.n, self.precision, int(part[0]), int(part[1]),
                                                   ctypes.byref(out)))
        self._h = out

    @staticmethod
    def handle_array(trees):
        return (ctypes.c_void_p * len(trees))(*[t._h for t in trees])

    def device_range(self, lo, hi):
        """(device pointer, bytes) of the packed blocks of trees [lo, hi)."""
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_int64()
        check(_lib.lib().pk_forest_device_range(self._h, int(lo), int(hi), ctypes.byref(ptr), ctypes.byref(nbytes)))
        return ptr.value or 0, nbytes.value

    def nbytes(self):
        b = ctypes.c_int64()
        check(_lib.lib().pk_forest_bytes(self._h, ctypes.byref(b)))
        return b.value

    def gram_stats(self, tile, part=0, nparts=1):
        c = (ctypes.c_int64 * 4)()
        check(_lib.lib().pk_forest_gram_stats(self._h, int(tile), int(part), int(nparts), c))
        return dict(dps=int(c[0]), node_pairs=int(c[1]), cells=int(c[2]), child_reads=int(c[3]))

    def pair_stats(self, other, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        c = (ctypes.c_int64 * 4)()
        check(_lib.lib().pk_forest_pair_stats(self._h, other._h, pairs.ctypes.data, pairs.shape[0], c))
        return dict(dps=int(c[0]), node_pairs=int(c[1]), cells=int(c[2]), child_reads=int(c[3]))

    def kernel_pairs(self, other, pairs, params):
        """K for (i in self, j in other) pairs: forests stay on the device, pair list and results are host arrays."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.empty(pairs.shape[0], np.float64)
        check(_lib.lib().pk_forest_kernel_pairs_host(self.ctx._h, self._h, other._h, pairs.ctypes.data, pairs.shape[0],
                                                     ctypes.byref(params), out.ctypes.data))
        return out

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value and self.ctx._h is not None and self.ctx.pid == os.getpid():
            _lib.lib().pk_forest_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CONTEXTS = {}


def get_context(device=None):
    """Per-process, per-device context; a forked child gets its own (lazily)."""
    if device is None:
        device = int(os.environ.get("PK_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = (os.getpid(), int(device))
    ctx = _CONTEXTS.get(key)
    if ctx is None:
        ctx = Context(device)
        _CONTEXTS[key] = ctx
    return ctx


def algorithmic_bytes(stats, precision=64):
    """w*(M+R) + 32*(n1+n2) + 8 per DP, summed; node_sum = sum over DPs of (n1 + n2)."""
    w = 8 if precision == 64 else 4
    return w * (stats["cells"] + stats["child_reads"]) + 32 * stats.get("node_sum", 0) + 8 * stats["dps"]


# ---------------------------------------------------------------------------------------------------
# the reference's class
# ---------------------------------------------------------------------------------------------------
class PhyloKernel(object):
    """Drop-in for ``phyloK2.PhyloKernel`` (constructor: phyloK2.py:10-62; same keyword names and defaults).

    How trees are prepared, stated once: tree OBJECTS and FlatTrees are always taken as the caller left them (the
    reference's ``kernel`` never re-prepares, phyloK2.py:166-172).  Newick STRINGS are parsed as written by ``kernel`` /
    ``kernel_batch`` (``prep=0, normalize='none'`` unless given), and prepared the way Kamphir prepares its trees
    (root length 0 -> ladderize -> ``self.normalize`` -> annotate, kamphir.py:279-282) by the entry points that stand
    for Kamphir's own loops: ``score_batch``, ``gram``, ``flatten_many``.  So for strings
    ``gram([a, b], normalised=False)[0, 1] == kernel_batch([a], [b], prep=PK_PREP_KAMPHIR, normalize=self.normalize)[0, 0]``.
    """

    def __init__(self,
                 kmat=None,
                 rotate='ladder',
                 rotate2='none',
                 subtree=False,
                 normalize='mean',
                 sigma=1,
                 gaussFactor=1,
                 withLengths=True,
                 decayFactor=0.1,
                 verbose=False,
                 resolve_poly=False,
                 **kwargs):
        self.trees = []
        self.kmat = []
        self.is_kmat_computed = False
        self.rotate = rotate
        self.rotate2 = rotate2
        self.subtree = subtree
        self.normalize = normalize
        self.sigma = sigma
        self.gaussFactor = gaussFactor
        self.withLengths = withLengths
        self.decayFactor = decayFactor
        self.verbose = verbose
        self.resolve_poly = resolve_poly
        self.pcache = {}
        self.subtrees = {}
        # additive knobs (not in the reference): compute precision, GPU index, tip-state count
        self.precision = int(kwargs.pop('precision', 64))
        self.device = kwargs.pop('device', None)
        self.n_states = int(kwargs.pop('n_states', 0))
        if self.precision not in (32, 64):
            raise ValueError("precision must be 64 (fp64, 1e-9) or 32 (fp32 cells, 1e-5)")
        if self.verbose:
            print('creating PhyloKernel with settings')
            print('sigma = {}'.format(self.sigma))
            print('gaussFactor = {}'.format(self.gaussFactor))
            print('decayFactor = {}'.format(self.decayFactor))

    @property
    def ntrees(self):
        return len(self.trees)

    # -- plumbing -------------------------------------------------------------------------------------
    def _ctx(self):
        return get_context(self.device)

    def _params(self):
        return pk_params(float(self.decayFactor), float(self.gaussFactor), float(self.sigma), self.precision, 0)

    def __getstate__(self):
        # dill/pickle safety (kamphir.py:27-32 pickles the bound method, i.e. the whole object): no raw pointers
        return dict(self.__dict__)

    # -- tree preparation, host side ---------------------------------------------------------------------
    def normalize_tree(self, t, mode='median'):
        """In-place rescale of a Bio.Phylo-like tree, phyloK2.py:101-130 (``//`` where py2 relied on int ``/``)."""
        branches = t.get_nonterminals() + t.get_terminals()
        nbranches = len(branches) - 1
        targets = branches[int(not t.rooted):]
        if mode == 'mean':
            tree_length = t.total_branch_length() - t.root.branch_length
            scale = tree_length / nbranches
        elif mode == 'median':
            lengths = sorted(b.branch_length for b in targets)
            if nbranches % 2 == 0:
                scale = (lengths[(nbranches // 2) - 1] + lengths[nbranches // 2]) / 2.
            else:
                scale = lengths[nbranches // 2]
        else:
            return
        for b in targets:
            b.branch_length /= scale
        _drop_cached_flat(t)

    def annotate_tree(self, t):
        """In-place annotations of phyloK2.py:132-146: production, index, bl, sqbl."""
        _drop_cached_flat(t)
        for tip in t.get_terminals():
            tip.production = 0
        for i, node in enumerate(t.get_nonterminals(order='postorder')):
            kids = node.clades
            node.production = sum(1 for c in kids if c.production == 0) + 1
            node.index = i
            node.bl = [c.branch_length for c in kids]
            node.sqbl = sum(b ** 2 for b in node.bl)

    def flatten(self, tree, prep=0, normalize='none'):
        """FlatTree from a FlatTree (returned as is), Newick text, or a Bio.Phylo-like tree object.

        For text, ``prep``/``normalize`` select the C++ flattener's kamphir-style preparation
        (``PK_PREP_KAMPHIR`` + ``'mean'`` = kamphir.py:279-282).  A tree *object* is taken as the caller left it
        (the reference's ``kernel`` never re-prepares, phyloK2.py:166-172): it is walked in post-order and its
        ``node.bl`` annotations (or, if absent, the children's branch lengths) are sent to the library.
        """
        if isinstance(tree, FlatTree):
            return tree
        if isinstance(tree, (str, bytes)):
            return FlatTree.from_newick(tree, prep, normalize, self.n_states)
        # A tree object is flattened once and the FlatTree is kept on it: the reference's kernel() reads the snapshot that
        # annotate_tree left on the nodes (node.bl, phyloK2.py:169-172,184-185), not the live branch lengths, so the
        # cached form stays valid until annotate_tree / normalize_tree run again (they drop it).  kamphir.py:290 passes
        # the same observed tree 2 * nreps times per MCMC step.
        cached = getattr(tree, '_pk_flat', None)
        if isinstance(cached, FlatTree) and cached._h is not None:
            return cached
        nodes = tree.get_nonterminals(order='postorder')
        if not nodes:
            raise ValueError("tree has no internal node (reference: IndexError at phyloK2.py:169)")
        if not hasattr(nodes[0], 'production'):          # phyloK2.py:169-172
            self.annotate_tree(tree)
        index = {id(nd): i for i, nd in enumerate(nodes)}
        cidx = np.empty((len(nodes), 2), np.int32)
        bl = np.empty((len(nodes), 2), np.float64)
        for i, nd in enumerate(nodes):
            kids = nd.clades
            if len(kids) != 2:
                raise ValueError("internal node %d has %d children; the kernel is defined for strictly binary trees "
                                 "(reference: IndexError at phyloK2.py:185-189)" % (i, len(kids)))
            lens = getattr(nd, 'bl', None)
            if lens is None:
                lens = [c.branch_length for c in kids]
            for s in (0, 1):
                cidx[i, s] = index.get(id(kids[s]), -1) if kids[s].clades else -1
                if lens[s] is None:
                    raise ValueError("branch without a length (reference: TypeError at phyloK2.py:146)")
                bl[i, s] = lens[s]
        flat = FlatTree.from_arrays(cidx, bl)
        try:
            tree._pk_flat = flat
        except AttributeError:            # an object with __slots__: nothing to hang the cache on
            pass
        return flat

    def flatten_list(self, trees, prep=0, normalize='none'):
        """flatten() for a sequence; runs of Newick strings go through one multi-threaded library call."""
        trees = list(trees)
        out = [None] * len(trees)
        text_idx = [i for i, t in enumerate(trees) if isinstance(t, (str, bytes))]
        if len(text_idx) > 1:
            flats = FlatTree.many_from_newick_list([trees[i] for i in text_idx], prep, normalize, self.n_states)
            for i, f in zip(text_idx, flats):
                out[i] = f
        for i, t in enumerate(trees):
            if out[i] is None:
                out[i] = self.flatten(t, prep, normalize)
        return out

    def flatten_many(self, newick_text, prep=PK_PREP_KAMPHIR, normalize=None, nthreads=0):
        """Multi-tree Newick text -> list of FlatTree, prepared like kamphir.py:279-282 by default."""
        return FlatTree.many_from_newick(newick_text, prep, self.normalize if normalize is None else normalize,
                                         self.n_states, nthreads)

    # -- the kernel ---------------------------------------------------------------------------------------
    def kernel(self, t1, t2, myrank=None, nprocs=None, output=None):
        """K(t1, t2), phyloK2.py:158-209, evaluated on the GPU.

        ``myrank``/``nprocs`` select the reference's row-striped mode, which is wrong by its own FIXME
        (phyloK2.py:223).  Here rank 0 (or None) returns the full K and every other rank 0.0, so the sum that
        ``kernel_parallel`` takes over ranks is the correct value.
        """
        if myrank is not None and nprocs and myrank % nprocs != 0:
            k = 0.0
        else:
            # single-tree forests stay on the device with their FlatTree (which stays with the tree object): a call on
            # trees seen before is one launch plus an 8-byte copy
            f1, f2 = self.flatten(t1), self.flatten(t2)
            ctx = self._ctx()
            d1 = f1.device_forest(ctx, self.precision)
            d2 = d1 if f2 is f1 else f2.device_forest(ctx, self.precision)
            k = float(d1.kernel_pairs(d2, _PAIR00, self._params())[0])
            if self.precision == 32 and not math.isfinite(k):      # fp32 cells overflowed: this pair again in fp64
                p = pk_params(float(self.decayFactor), float(self.gaussFactor), float(self.sigma), 64, 0)
                k = float(f1.device_forest(ctx, 64).kernel_pairs(f2.device_forest(ctx, 64), _PAIR00, p)[0])
        if output is None:
            return k
        output.put(k)

    def kernel_parallel(self, t1, t2, nthreads):
        """phyloK2.py:211-236 without the processes: one GPU launch already uses the whole device."""
        return self.kernel(t1, t2)

    def kernel_batch(self, trees_a, trees_b=None, pairs=None, prep=0, normalize='none'):
        """K for many pairs in one launch.  ``pairs`` = [(ia, ib), ...] -> vector; None -> len(a) x len(b) matrix.

        Like ``kernel``, trees are taken as the caller prepared them: Newick strings are parsed with ``prep`` /
        ``normalize`` (default: as written, no rescaling).  Pass ``prep=PK_PREP_KAMPHIR, normalize=self.normalize`` for
        the preparation that ``gram`` and ``score_batch`` apply to strings (kamphir.py:279-282)."""
        fa = self.flatten_list(trees_a, prep, normalize)
        same = trees_b is None
        fb = fa if same else self.flatten_list(trees_b, prep, normalize)
        if pairs is None:
            pr = np.stack(np.meshgrid(np.arange(len(fa)), np.arange(len(fb)), indexing='ij'), -1).reshape(-1, 2)
        else:
            pr = np.asarray(pairs).reshape(-1, 2)
        pr = np.ascontiguousarray(pr, dtype=np.int32)
        out = np.empty(pr.shape[0], np.float64)
        if pr.shape[0] == 0:
            return out
        arr_a = (ctypes.c_void_p * len(fa))(*[t._h for t in fa])
        arr_b = arr_a if same else (ctypes.c_void_p * len(fb))(*[t._h for t in fb])
        p = self._params()
        check(_lib.lib().pk_kernel_pairs(self._ctx()._h, arr_a, len(fa), arr_b, len(fb), pr.ctypes.data, pr.shape[0],
                                         ctypes.byref(p), out.ctypes.data))
        return out if pairs is not None else out.reshape(len(fa), len(fb))

    def score_batch(self, obs, sims, ref_denom=None, kadj=None, use_kcoal=False):
        """One ABC proposal's scores (kamphir.py:288-306, 434-450) for a target tree and its simulated replicates.

        Returns a dict with ``k``, ``denom``, ``knorm`` (arrays), ``ref_denom`` and, with ``use_kcoal``,
        ``kcoal`` and ``score = knorm * kcoal**kadj`` and their mean (``mean_score``).
        """
        fo = self.flatten(obs, PK_PREP_KAMPHIR, self.normalize)
        sims = list(sims)
        if sims and all(isinstance(t, (str, bytes)) for t in sims):
            # Newick text straight from the simulators (kamphir.py:318): parse, prep, DP launch, kcoal and the scores in
            # one library call -- no per-tree Python objects
            raw = [t.encode("utf-8") if isinstance(t, str) else t for t in sims]
            n = len(raw)
            ptrs = (ctypes.c_char_p * n)(*raw)
            lens = (ctypes.c_size_t * n)(*[len(b) for b in raw])
            out = np.empty((5, n))
            ref = ctypes.c_double(float('nan') if ref_denom is None else float(ref_denom))
            mean = ctypes.c_double()
            p = self._params()
            check(_lib.lib().pk_score_newick_batch(
                self._ctx()._h, fo._h, ptrs, lens, n, int(PK_PREP_KAMPHIR), PK_NORM[self.normalize], int(self.n_states), 0,
                ctypes.byref(p), ctypes.byref(ref), (1.0 if kadj is None else float(kadj)) if use_kcoal else float('nan'),
                out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, out[3].ctypes.data, out[4].ctypes.data,
                ctypes.byref(mean)))
            res = dict(k=out[0], denom=out[1], knorm=out[2], ref_denom=ref.value, mean_score=mean.value)
            if use_kcoal:
                res['kcoal'] = out[3]
                res['score'] = out[4]
            self._check_log_domain(res)
            return res
        fs = self.flatten_list(sims, PK_PREP_KAMPHIR, self.normalize)
        n = len(fs)
        k, d, kh = np.empty(n), np.empty(n), np.empty(n)
        ref = ctypes.c_double(float('nan') if ref_denom is None else float(ref_denom))
        arr = (ctypes.c_void_p * n)(*[t._h for t in fs])
        p = self._params()
        check(_lib.lib().pk_score_batch(self._ctx()._h, fo._h, arr, n, ctypes.byref(p), ctypes.byref(ref),
                                        k.ctypes.data, d.ctypes.data, kh.ctypes.data))
        res = dict(k=k, denom=d, knorm=kh, ref_denom=ref.value)
        self._check_log_domain(res)
        if use_kcoal:
            kc = np.empty(n)
            check(_lib.lib().pk_kcoal_batch(arr, n, fo._h, kc.ctypes.data))
            res['kcoal'] = kc
            res['score'] = kh * np.power(kc, 1.0 if kadj is None else kadj)
            self._check_log_domain(res)
            res['mean_score'] = float(res['score'].sum() / n)       # kamphir.py:450
        else:
            res['mean_score'] = float(kh.sum() / n)
        return res

    @staticmethod
    def _check_log_domain(res):
        """kamphir.py:300 takes math.log of k, ref_denom and tree_denom: a replicate without a single matched node pair
        (decayFactor 0, or labelled trees that share no production class) is a ValueError there, and here."""
        bad = np.flatnonzero(~((res['k'] > 0) & (res['denom'] > 0)))
        if bad.size or not res['ref_denom'] > 0:
            raise ValueError("math domain error (kamphir.py:300): K = 0 for %s"
                             % ("replicate %d" % bad[0] if bad.size else "the observed tree"))
        for key in ('knorm', 'score'):          # a NaN score would silently accept or reject an MCMC step
            if key in res and not np.all(np.isfinite(res[key])):
                raise ValueError("non-finite %s for replicate %d" % (key, int(np.flatnonzero(~np.isfinite(res[key]))[0])))

    def score_batch_sharded(self, obs, sims, rank, world, group=None, ref_denom=None, kadj=None, use_kcoal=False):
        """score_batch over the ranks of a torch.distributed group (one process per GPU): rank r scores every
        world-th replicate, one all-gather returns the full result to every rank (kamphir_b200/score.py)."""
        import torch
        from .score import sharded_scores
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.distributed.get_backend(group) == "nccl" else None
        return sharded_scores(lambda part: self.score_batch(obs, part, ref_denom=ref_denom, kadj=kadj, use_kcoal=use_kcoal),
                              sims, rank, world, group, dev)

    def gram(self, trees, normalised=True):
        """All-pairs matrix over ``trees`` (FlatTree / Newick / tree objects): raw K or K/sqrt(K_ii K_jj)."""
        ft = self.flatten_list(trees, PK_PREP_KAMPHIR, self.normalize)
        n = len(ft)
        out = np.empty((n, n), np.float64)
        arr = (ctypes.c_void_p * n)(*[t._h for t in ft])
        p = self._params()
        check(_lib.lib().pk_gram(self._ctx()._h, arr, n, ctypes.byref(p), 1 if normalised else 0, out.ctypes.data))
        return out

    # -- the reference's (broken) matrix surface, made to work: phyloK2.py:68-99, 148-156 -------------------
    def load_trees_from_file(self, handle):
        """Parse a Newick file into ``self.trees`` (FlatTrees): ladderize when ``rotate == 'ladder'``, normalise with
        ``self.normalize``, annotate (phyloK2.py:73-93).  The reference's other rotate/gravitate/polytomy options
        call functions that do not exist (phyloK2.py:79-91) and are rejected here."""
        if self.rotate not in ('ladder', 'none', None) or self.rotate2 != 'none' or self.resolve_poly:
            raise NotImplementedError("only rotate='ladder'/'none' is defined (phyloK2.py:79-91 reference undefined names)")
        text = handle.read() if hasattr(handle, 'read') else open(handle).read()
        prep = PK_PREP_LADDERIZE if self.rotate == 'ladder' else 0
        norm = self.normalize if self.normalize != 'none' else 'none'
        if norm != 'none':
            prep |= PK_PREP_ZERO_ROOT          # root length is "meaningless" (phyloK2.py:105) and None breaks :113
        self.trees = FlatTree.many_from_newick(text, prep, norm, self.n_states)
        self.kmat = np.zeros((self.ntrees, self.ntrees))
        self.is_kmat_computed = False
        self.delta_values = {}

    def compute_matrix(self):
        """kmat[i,j] = kmat[j,i] = kernel(trees[i], trees[j]) for j >= i, raw K (phyloK2.py:148-156)."""
        self.kmat = self.gram(self.trees, normalised=False)
        if self.verbose:
            for i in range(self.ntrees):
                for j in range(i, self.ntrees):
                    print('%d\t%d\t%f' % (i, j, self.kmat[i, j]))
        self.is_kmat_computed = True

    # -- consumers of the kernel matrix (the uses README.md:3-5 of the reference names: kernel-PCA, kernel-SVM) -----------
    def normalised_matrix(self):
        """K[i,j] / sqrt(K[i,i] K[j,j]) of the computed matrix (the normalisation of kamphir.py:300 applied to ``kmat``)."""
        if not getattr(self, 'is_kmat_computed', False):
            raise ValueError("compute_matrix() has not run")
        d = np.sqrt(np.diag(self.kmat))
        return self.kmat / np.outer(d, d)

    def save_matrix(self, path, normalised=False):
        """Write the kernel matrix as ``.npy`` (the input format of scikit-learn's ``kernel='precomputed'`` estimators)."""
        np.save(path, self.normalised_matrix() if normalised else self.kmat)

    def kernel_pca(self, n_components=2, normalised=True):
        """Kernel principal components of the tree set from the computed matrix: the matrix is double-centred,
        K_c = (I - 1/N) K (I - 1/N), and tree i gets the coordinates sqrt(w_k) v_k[i] of the ``n_components`` largest
        eigenpairs (w_k, v_k) of K_c.  Returns (coordinates [N, n_components], fraction of the variance per component).
        Host-side numpy on the N x N result; the DP work was done by ``compute_matrix`` / ``gram``."""
        k = self.normalised_matrix() if normalised else np.asarray(self.kmat, dtype=np.float64)
        n = k.shape[0]
        if not 1 <= n_components <= n:
            raise ValueError("n_components must be between 1 and the number of trees")
        rm = k.mean(axis=0)
        kc = k - rm[None, :] - rm[:, None] + rm.mean()
        w, v = np.linalg.eigh((kc + kc.T) * 0.5)
        order = np.argsort(w)[::-1][:n_components]
        w_top = np.clip(w[order], 0.0, None)
        total = np.clip(w, 0.0, None).sum()
        coords = v[:, order] * np.sqrt(w_top)[None, :]
        # sign convention: the entry of largest magnitude of every component is positive (eigenvectors have no sign)
        for c in range(coords.shape[1]):
            if coords[np.argmax(np.abs(coords[:, c])), c] < 0:
                coords[:, c] = -coords[:, c]
        return coords, (w_top / total if total > 0 else np.zeros_like(w_top))


_PAIR00 = np.zeros((1, 2), np.int32)


def _drop_cached_flat(tree):
    try:
        if getattr(tree, '_pk_flat', None) is not None:
            tree._pk_flat = None
    except AttributeError:
        pass


def knorm(k, ref_denom, tree_denom):
    """kamphir.py:300 (host scalar form)."""
    return math.exp(math.log(k) - 0.5 * (math.log(ref_denom) + math.log(tree_denom)))
