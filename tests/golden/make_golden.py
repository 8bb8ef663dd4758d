#!/usr/bin/env python
"""Generate tests/golden/kernel_kat.json by running the REFERENCE's own, unmodified phyloK2.py.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

``/root/reference/phyloK2.py`` is imported as-is; its single third-party import (``from Bio import Phylo``) is
satisfied by the stand-in in ``oracle/biophylo_shim`` (Biopython is not installed and there is no network).
Every expected value in the JSON is a float returned by ``phyloK2.PhyloKernel.kernel`` (or produced by its
``normalize_tree``), stored with ``repr`` precision.  The Newick inputs are either files copied verbatim
(data, not source) into ``tests/golden/trees/`` or inline strings stored in the JSON.

Tree preparation modes:
  raw            parse only (``tests/unittests.py:93-101``)
  kamphir:mean   root bl 0 -> ladderize -> normalize_tree('mean') -> annotate_tree (``kamphir.py:153-156,279-282``)
  ladder:none    ladderize only
"""
import json
import math
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REPO, "oracle", "biophylo_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from io import StringIO  # noqa: E402
from Bio import Phylo  # noqa: E402  (stand-in)
import phyloK2  # noqa: E402  (the reference, unmodified)
from kamphir_b200 import synth  # noqa: E402

FILES = {
    "test.tree": "tests/test.tree",
    "SIRTree.n100.nwk": "projects/hivepi/data/SIRTree.n100.nwk",
    "SIRTree.n300.nwk": "projects/hivepi/data/SIRTree.n300.nwk",
    "SIRTree.n1000.nwk": "projects/hivepi/data/SIRTree.n1000.nwk",
    "cn.crf07.timetree.nwk": "projects/hivepi/data/cn.crf07.gp120.RLRootToTip.timetree.nwk",
    "hivSimulation.nwk": "rcolgem/pkg/inst/extdata/hivSimulation.nwk",
    "sirModel0.nwk": "rcolgem/pkg/inst/extdata/sirModel0.nwk",
}

T1 = "( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;"      # tests/unittests.py:87
T2 = "( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;"    # tests/unittests.py:88


def load(spec):
    if "file" in spec:
        trees = list(Phylo.parse(os.path.join(HERE, "trees", spec["file"]), "newick"))
        return trees[spec.get("index", 0)]
    return Phylo.read(StringIO(spec["newick"]), "newick")


def prepare(pk, tree, prep):
    if prep == "raw":
        return tree
    kind, mode = prep.split(":")
    if kind == "kamphir":
        tree.root.branch_length = 0.
        tree.ladderize()
        pk.normalize_tree(tree, mode)
        pk.annotate_tree(tree)
    elif kind == "ladder":
        tree.ladderize()
    else:
        raise ValueError(prep)
    return tree


def main():
    os.makedirs(os.path.join(HERE, "trees"), exist_ok=True)
    for dst, src in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, "trees", dst))
        os.chmod(os.path.join(HERE, "trees", dst), 0o644)

    F = lambda name, index=0: {"file": name, "index": index}  # noqa: E731
    S = lambda text: {"newick": text}  # noqa: E731
    kam = dict(lam=0.2, tau=2.0, sigma=1)
    unit = dict(lam=0.5, tau=1, sigma=1)
    cases = [
        dict(name="unittest_kernel", a=S(T1), b=S(T2), prep="raw", **unit,
             ref_literal="1.125*(1+exp(-0.0625))  tests/unittests.py:95"),
        dict(name="unittest_kernel_large", a=F("test.tree", 0), b=F("test.tree", 1), prep="raw", **unit,
             ref_literal="3042.58677467 (7 places)  tests/unittests.py:101"),
        dict(name="testtree_01_kamphir", a=F("test.tree", 0), b=F("test.tree", 1), prep="kamphir:mean", **kam),
        dict(name="testtree_00_kamphir", a=F("test.tree", 0), b=F("test.tree", 0), prep="kamphir:mean", **kam),
        dict(name="testtree_11_kamphir", a=F("test.tree", 1), b=F("test.tree", 1), prep="kamphir:mean", **kam),
        dict(name="testtree_00_defaults", a=F("test.tree", 0), b=F("test.tree", 0), prep="kamphir:mean",
             lam=0.1, tau=1, sigma=1),
        dict(name="testtree_01_sigma", a=F("test.tree", 0), b=F("test.tree", 1), prep="kamphir:mean",
             lam=0.3, tau=0.7, sigma=0.25),
        dict(name="testtree_01_ladder_only", a=F("test.tree", 0), b=F("test.tree", 1), prep="ladder:none", **kam),
    ]
    for f in ("SIRTree.n100.nwk", "SIRTree.n300.nwk", "SIRTree.n1000.nwk", "cn.crf07.timetree.nwk",
              "hivSimulation.nwk", "sirModel0.nwk"):
        cases.append(dict(name="self_" + f, a=F(f), b=F(f), prep="kamphir:mean", **kam))
    cases += [
        dict(name="n100_vs_n300", a=F("SIRTree.n100.nwk"), b=F("SIRTree.n300.nwk"), prep="kamphir:mean", **kam),
        dict(name="n300_vs_n100", a=F("SIRTree.n300.nwk"), b=F("SIRTree.n100.nwk"), prep="kamphir:mean", **kam),
        dict(name="sir0_vs_hivsim", a=F("sirModel0.nwk"), b=F("hivSimulation.nwk"), prep="kamphir:mean", **kam),
        dict(name="crf07_vs_n300_raw", a=F("cn.crf07.timetree.nwk"), b=F("SIRTree.n300.nwk"), prep="raw",
             lam=0.2, tau=5000.0, sigma=1),
        dict(name="testtree0_vs_n1000", a=F("test.tree", 0), b=F("SIRTree.n1000.nwk"), prep="kamphir:mean", **kam),
    ]
    # seeded synthetic trees (inline so the fixture does not depend on the generator staying unchanged)
    syn = {
        "king50": synth.kingman_newick(50, 11), "king51": synth.kingman_newick(51, 12),
        "king100a": synth.kingman_newick(100, 1000), "king100b": synth.kingman_newick(100, 1001),
        "king200a": synth.kingman_newick(200, 2000), "king200b": synth.kingman_newick(200, 2001),
        "yule100": synth.yule_newick(100, 21), "yule200": synth.yule_newick(200, 22),
        "king500": synth.kingman_newick(500, 3000), "yule500": synth.yule_newick(500, 3001),
        "cat40": synth.caterpillar_newick(40, 5), "cat41": synth.caterpillar_newick(41, 6),
        "bal6": synth.balanced_newick(6, 7), "bal5": synth.balanced_newick(5, 8),
        "cherry": "(a:0.3,b:0.7);", "cherry2": "(a:1.3,b:0.2);", "three": "((a:0.1,b:0.2):0.3,c:0.4);",
    }
    pairs = [("king50", "king51"), ("king100a", "king100b"), ("king100a", "yule100"), ("king200a", "king200b"),
             ("king200a", "yule200"), ("yule200", "yule200"), ("king500", "yule500"), ("king500", "king500"),
             ("cat40", "cat41"), ("cat40", "bal6"), ("bal6", "bal5"), ("bal6", "bal6"), ("cherry", "cherry2"),
             ("cherry", "three"), ("three", "three"), ("cherry", "king50"), ("king100a", "cat40")]
    for a, b in pairs:
        cases.append(dict(name="syn_%s_%s" % (a, b), a=S(syn[a]), b=S(syn[b]), prep="kamphir:mean", **kam))
    for a, b in [("king50", "king51"), ("cat40", "bal6"), ("king100a", "yule100")]:
        cases.append(dict(name="synraw_%s_%s" % (a, b), a=S(syn[a]), b=S(syn[b]), prep="raw",
                          lam=0.35, tau=1.5, sigma=1))

    for c in cases:
        pk = phyloK2.PhyloKernel(decayFactor=c["lam"], gaussFactor=c["tau"], sigma=c["sigma"])
        ta = prepare(pk, load(c["a"]), c["prep"])
        tb = prepare(pk, load(c["b"]), c["prep"])
        c["K"] = pk.kernel(ta, tb)
        c["n_internal"] = [len(ta.get_nonterminals()), len(tb.get_nonterminals())]
        print("%-36s K = %r" % (c["name"], c["K"]))

    # the reference's own literals must hold before anything is written
    assert cases[0]["K"] == 1.125 * (1 + math.exp(-0.0625)), cases[0]["K"]
    assert round(cases[1]["K"] - 3042.58677467, 7) == 0, cases[1]["K"]

    # normalize_tree golden (tests/unittests.py:109-113): T1 mean-normalised = every non-root branch / 0.375.
    # (Newer Biopython leaves the root length None -> TypeError at phyloK2.py:113; kamphir always sets it to 0.)
    pk = phyloK2.PhyloKernel()
    t = Phylo.read(StringIO(T1), "newick")
    t.root.branch_length = 0.
    expected = [c.branch_length / 0.375 for c in t.find_clades(order="postorder")][:-1]
    pk.normalize_tree(t, "mean")
    got = [c.branch_length for c in t.find_clades(order="postorder")][:-1]
    assert got == expected
    norm = dict(newick=T1, mode="mean", postorder_non_root=got, divisor=0.375)

    with open(os.path.join(HERE, "kernel_kat.json"), "w") as fh:
        json.dump(dict(generator="tests/golden/make_golden.py", reference="/root/reference/phyloK2.py (unmodified)",
                       python=sys.version.split()[0], cases=cases, normalize=norm), fh, indent=1)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
