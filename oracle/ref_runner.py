"""Runs the REFERENCE's own ``phyloK2.py`` (unmodified) as the CPU baseline  --  TEST INFRASTRUCTURE ONLY.

``oracle/Makefile`` places ``/root/reference/phyloK2.py`` into ``oracle/_ref/`` at build time (git-ignored; it travels to
the GPU box with the push like a built ``.so``).  Its one third-party import (``from Bio import Phylo``) is satisfied
by the stand-in in ``oracle/biophylo_shim`` (Biopython is absent from the image and un-pinned in the reference,
``README.md:15``).  The module is run under the image's CPython 3 because the reference's Python 2.7 does not exist here;
``phyloK2.py`` has no py2-only syntax.

What is timed is exactly ``PhyloKernel.kernel(t1, t2)`` (``phyloK2.py:158-209``) on trees prepared the way Kamphir
prepares them (``kamphir.py:153-156, 279-282``: root length 0 -> ladderize -> normalize_tree -> annotate_tree).

Only ``bench.py`` (cpu_baseline / --impl reference) and ``tests/`` may import this module.
"""
import importlib.util
import os
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "biophylo_shim")
REF_PATH = os.path.join(_HERE, "_ref", "phyloK2.py")
_MOD = None


def available():
    return os.path.exists(REF_PATH)


def module():
    """The reference module, imported from oracle/_ref/phyloK2.py (raises if the recipe has not been run)."""
    global _MOD
    if _MOD is None:
        if not available():
            raise RuntimeError("oracle/_ref/phyloK2.py is missing: run `make -C oracle` where /root/reference exists")
        if _SHIM not in sys.path:
            sys.path.insert(0, _SHIM)
        spec = importlib.util.spec_from_file_location("phyloK2_reference", REF_PATH)
        mod = importlib.util.module_from_spec(spec)
        dont = sys.dont_write_bytecode
        sys.dont_write_bytecode = True
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = dont
        _MOD = mod
    return _MOD


def make_kernel(decayFactor=0.2, gaussFactor=2.0, sigma=1.0, normalize="mean"):
    return module().PhyloKernel(decayFactor=decayFactor, gaussFactor=gaussFactor, sigma=sigma, normalize=normalize)


def prepare(pk, newick, normalize="mean"):
    """kamphir.py:279-282 on a Newick string, with the reference's own normalize_tree / annotate_tree."""
    from io import StringIO
    from Bio import Phylo      # the stand-in
    tree = Phylo.read(StringIO(newick), "newick")
    tree.root.branch_length = 0.
    tree.ladderize()
    pk.normalize_tree(tree, normalize)
    pk.annotate_tree(tree)
    return tree


def time_pairs(pk, trees, pairs):
    """[K], seconds: kernel() calls only (prep excluded), one process."""
    t0 = time.perf_counter()
    ks = [pk.kernel(trees[a], trees[b]) for a, b in pairs]
    return ks, time.perf_counter() - t0
