"""Round-2 GPU tests: the reference-facing call on tree objects (cache + device-resident trees), fp32 overflow safety,
BASELINE.json configs[0] and configs[2] as stated, the cost-weighted Gram partition, trees beyond the shared-memory
staging limit.  Needs a B200:  pytest -m gpu.  Nothing here reads /root/reference.
"""
import math
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN
from kamphir_b200 import FlatTree, Forest, PhyloKernel, get_context, synth
from kamphir_b200._lib import pk_params
from oracle import c_oracle
from oracle import phylok_oracle as po

pytestmark = pytest.mark.gpu
TOL = {64: 1e-9, 32: 1e-5}
KAM = dict(decayFactor=0.2, gaussFactor=2.0)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def oracle_flat(nwk, prep="kamphir:mean"):
    return c_oracle.CFlat(po.prepare(po.read_newick(nwk), prep))


def kamphir_prep(pk, nwk, mode="mean"):
    """kamphir.py:279-282 on a Bio.Phylo-like object with the class's own methods."""
    t = po.read_newick(nwk)
    t.root.branch_length = 0.
    t.ladderize()
    pk.normalize_tree(t, mode)
    pk.annotate_tree(t)
    return t


@pytest.fixture(autouse=True)
def _default_config():
    get_context().configure(0, 0, 0)
    yield
    get_context().configure(0, 0, 0)


# ---------------------------------------------------------------------------------------------------
# the call kamphir.py makes: kernel(tree object, tree object)
# ---------------------------------------------------------------------------------------------------
def test_kernel_on_tree_objects_is_cached_and_follows_re_preparation():
    pk = PhyloKernel(**KAM)
    a, b = synth.kingman_newick(60, 1), synth.yule_newick(75, 2)
    ta, tb = kamphir_prep(pk, a), kamphir_prep(pk, b)
    fa, fb = oracle_flat(a), oracle_flat(b)
    want = c_oracle.kernel(fa, fb, 0.2, 2.0)
    launches0 = get_context().last_stats()["launches"]
    k1 = pk.kernel(ta, tb)                      # kamphir.py:290
    k2 = pk.kernel(ta, tb)                      # again: no walk, no upload
    assert k1 == k2 and abs(k1 - want) <= 1e-9 * want
    assert get_context().last_stats()["launches"] == launches0 + 2
    flat_a = ta._pk_flat
    assert flat_a is pk.flatten(ta) and flat_a._dev
    assert abs(pk.kernel(tb, tb) - c_oracle.kernel(fb, fb, 0.2, 2.0)) <= 1e-9 * want      # kamphir.py:291
    # a tree that is prepared again is flattened and uploaded again
    pk.normalize_tree(ta, "median")
    pk.annotate_tree(ta)
    k3 = pk.kernel(ta, tb)
    assert ta._pk_flat is not flat_a
    fm = c_oracle.CFlat(po.annotate(ta))
    assert abs(k3 - c_oracle.kernel(fm, fb, 0.2, 2.0)) <= 1e-9 * k3 and k3 != k1
    # strings and FlatTrees through the same call
    assert abs(pk.kernel(FlatTree.from_newick(a), FlatTree.from_newick(b)) - want) <= 1e-9 * want


def test_reference_literals_through_the_gpu_path():
    """tests/unittests.py:95 and :101, the two values the reference's own tests assert."""
    pk = PhyloKernel(decayFactor=0.5, gaussFactor=1)
    t1 = po.read_newick("( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;")
    t2 = po.read_newick("( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;")
    assert abs(pk.kernel(t1, t2) - 1.125 * (1 + math.exp(-0.0625))) < 1e-14
    trees = po.parse_newick_file(os.path.join(GOLDEN, "trees", "test.tree"))
    assert round(pk.kernel(trees[0], trees[1]) - 3042.58677467, 7) == 0


# ---------------------------------------------------------------------------------------------------
# fp32 safety
# ---------------------------------------------------------------------------------------------------
def test_fp32_overflow_never_reaches_the_caller():
    """decayFactor 0.5 (the reference's own unit-test setting, tests/unittests.py:93) on a balanced 1024-tip tree against
    itself: C grows like x -> lambda (1 + x)^2 per level and passes 3e38 at depth 9 (K = 3.7e65).  Every host entry
    point answers with the fp64 value instead of inf / NaN.  (An 800-tip caterpillar stays small: its cells converge to
    lambda (sigma + lambda) (1 + C) = 3.)"""
    cat = synth.balanced_newick(10, 1)
    f = FlatTree.from_newick(cat)
    o = oracle_flat(cat)
    want = c_oracle.kernel(o, o, 0.5, 1.0)
    assert math.isfinite(want) and want > 1e38                      # beyond fp32, inside fp64
    deep = synth.caterpillar_newick(800, 3)
    od = oracle_flat(deep)
    kd = PhyloKernel(decayFactor=0.5, gaussFactor=1.0, precision=32).kernel(FlatTree.from_newick(deep), FlatTree.from_newick(deep))
    assert abs(kd - c_oracle.kernel(od, od, 0.5, 1.0)) <= 1e-5 * kd
    pk32 = PhyloKernel(decayFactor=0.5, gaussFactor=1.0, precision=32)
    k = pk32.kernel(f, f)
    assert math.isfinite(k) and abs(k - want) <= 1e-9 * want        # re-evaluated in fp64
    kb = pk32.kernel_batch([f], [f], pairs=[(0, 0)])[0]
    assert math.isfinite(kb) and abs(kb - want) <= 1e-9 * want
    small = FlatTree.from_newick(synth.kingman_newick(30, 5))
    mixed = pk32.kernel_batch([f, small], pairs=[(0, 0), (1, 1), (0, 1)])
    os_ = oracle_flat(synth.kingman_newick(30, 5))
    assert np.all(np.isfinite(mixed))
    assert abs(mixed[1] - c_oracle.kernel(os_, os_, 0.5, 1.0)) <= 1e-5 * mixed[1]          # untouched pairs stay fp32
    g = pk32.gram([f, small], normalised=False)
    assert np.all(np.isfinite(g)) and abs(g[0, 0] - want) <= 1e-9 * want
    res = pk32.score_batch(f, [f, f])
    assert np.all(np.isfinite(res["knorm"])) and abs(res["knorm"][0] - 1.0) < 1e-9
    # the reference's own lambda = 0.5 unit-test trees in fp32: no overflow, fp32 tolerance
    t1 = "( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;"
    t2 = "( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;"
    assert abs(pk32.kernel(t1, t2) - 1.125 * (1 + math.exp(-0.0625))) < 1e-5
    text = open(os.path.join(GOLDEN, "trees", "test.tree")).read()
    trees = po._split_trees_text(text)
    assert abs(pk32.kernel(trees[0], trees[1]) - 3042.58677467) < 1e-5 * 3042.6


def test_device_resident_gram_reports_fp32_overflow():
    from kamphir_b200.gram import GramRunner
    cat = synth.balanced_newick(10, 1)
    flats = [FlatTree.from_newick(cat), FlatTree.from_newick(synth.kingman_newick(40, 1))]
    r = GramRunner(flats, 0.5, 1.0, 1.0, 32, False, 1)
    r.upload()
    r.launch()
    r.finish()
    with pytest.raises(OverflowError):
        r.fetch()
    r.close()
    r = GramRunner(flats, 0.5, 1.0, 1.0, 64, False, 1)
    r.upload()
    r.launch()
    r.finish()
    assert np.all(np.isfinite(r.fetch()))
    r.close()


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: tests/test.tree[0] against 100 synthetic 100-tip coalescent trees
# ---------------------------------------------------------------------------------------------------
def test_config0_as_stated_test_tree_against_100_coalescent_trees():
    obs = po._split_trees_text(open(os.path.join(GOLDEN, "trees", "test.tree")).read())[0]
    sims = [synth.kingman_newick(100, i) for i in range(100)]           # seed = config index * 1000 + tree index
    pk = PhyloKernel(normalize="mean", **KAM)
    res = pk.score_batch(obs, [s.encode() for s in sims], use_kcoal=True)
    fo = oracle_flat(obs)
    ref = c_oracle.kernel(fo, fo, 0.2, 2.0)
    assert abs(ref - 577.30924562105) < 1e-9 * ref                      # SURVEY.md 8c KAT produced by the reference
    assert abs(res["ref_denom"] - ref) <= 1e-9 * ref
    for s in range(100):
        fs = oracle_flat(sims[s])
        k, d = c_oracle.kernel(fo, fs, 0.2, 2.0), c_oracle.kernel(fs, fs, 0.2, 2.0)
        assert abs(res["k"][s] - k) <= 1e-9 * k and abs(res["denom"][s] - d) <= 1e-9 * d
        assert abs(res["knorm"][s] - k / math.sqrt(ref * d)) <= 1e-9
    mean, scores = po.kamphir_evaluate(sims[:10], obs, decayFactor=0.2, gaussFactor=2.0, normalize="mean")
    assert rel(res["score"][:10], scores) < 1e-9
    # the same proposal through the reference-shaped calls on tree objects (kamphir.py:279-291)
    t_obs = kamphir_prep(pk, obs)
    for s in range(4):
        t = kamphir_prep(pk, sims[s])
        assert abs(pk.kernel(t_obs, t) - res["k"][s]) <= 1e-12 * res["k"][s]
        assert abs(pk.kernel(t, t) - res["denom"][s]) <= 1e-12 * res["denom"][s]


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: 500-tip multi-type trees, state matching on and off
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,prev", [(2, [0.5, 0.5]), (3, [0.2, 0.5, 0.3])])
def test_config2_at_its_stated_size(S, prev):
    """500 tips, S = 2 (DiffRisk, p = 0.5: settings/example.DiffRisk.json:20-28) and S = 3 (Stages), tip names
    "<i>_<deme>" with rcolgem's 1-based demes (rcolgem.py:274), fp64 and fp32, matching on and off.  The checker of the
    labelled mode is the build's own extension oracle (C restatement with explicit classes), itself checked against the
    recursive definition in tests/test_oracle.py: parity of this mode is unpinned by the reference (it has no such mode)."""
    nwk0 = [synth.kingman_newick(500, 2000 + i, states=S, prevalence=prev) for i in range(3)] + \
           [synth.yule_newick(500, 2100, states=S, prevalence=prev)]
    nwk = [re.sub(r"_(\d+)", lambda m: "_%d" % (int(m.group(1)) + 1), s) for s in nwk0]      # demes 1 .. S
    of = [po.prep_kamphir(po.read_newick(s), "mean", po.state_from_name) for s in nwk]
    lab = [c_oracle.labelled_cflat(f, S) for f in of]
    plain = [c_oracle.CFlat(f) for f in of]
    want_on = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in lab] for a in lab])
    want_off = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in plain] for a in plain])
    assert np.all(want_on < want_off)
    # the C restatement with explicit classes is the Python extension oracle (one 500-tip pair: ~1 s)
    assert abs(po.kernel_arrays_labelled(of[0], of[3], 0.2, 2.0) - want_on[0, 3]) <= 1e-12 * want_on[0, 3]
    for precision in (64, 32):
        on = PhyloKernel(n_states=S, precision=precision, **KAM)
        off = PhyloKernel(precision=precision, **KAM)
        f_on = [FlatTree.from_newick(s, 3, "mean", S) for s in nwk]
        f_off = [FlatTree.from_newick(s, 3, "mean", 0) for s in nwk]
        assert f_on[0].n_classes == 1 + 2 * S + S * S
        assert rel(on.kernel_batch(f_on), want_on) <= TOL[precision]
        assert rel(off.kernel_batch(f_off), want_off) <= TOL[precision]
        assert rel(on.gram(f_on, normalised=False), want_on) <= TOL[precision]
        # 0-based and 1-based deme numbers give the same classes up to renaming, hence the same kernel
        f0 = [FlatTree.from_newick(s, 3, "mean", S) for s in nwk0]
        assert rel(on.kernel_batch(f0), want_on) <= TOL[precision]
        # a proposal: labelled observed tree against labelled replicates
        res = on.score_batch(nwk[0], nwk[1:])
        wk = want_on[0, 1:] / np.sqrt(want_on[0, 0] * np.diag(want_on)[1:])
        assert rel(res["knorm"], wk) <= 10 * TOL[precision]


# ---------------------------------------------------------------------------------------------------
# Gram partition by cost
# ---------------------------------------------------------------------------------------------------
def test_cost_weighted_gram_parts_cover_the_triangle_and_balance_ragged_forests():
    from kamphir_b200.gram import GramRunner
    sizes = [20, 25, 30, 200, 220, 35, 40, 180, 22, 28, 33, 150, 45, 50, 21, 160, 24, 26, 210, 31, 29, 27]
    nwk = [synth.kingman_newick(n, 500 + i) for i, n in enumerate(sizes)]
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    want = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in orac] for a in orac])
    n, tile, world = len(flats), 4, 3
    full = np.zeros((n, n))
    seen = set()
    loads = []
    for r in range(world):
        run = GramRunner(flats, 0.2, 2.0, 1.0, 64, False, tile, "local", 0, 1)
        run.upload()
        run.rank, run.world = r, world          # evaluate part r of 3 on this one GPU into its own zeroed matrix
        tiles = [tuple(t) for t in run.my_tiles()]
        assert not (seen & set(tiles))
        seen |= set(tiles)
        st = run.my_stats()
        loads.append(st["node_pairs"])
        run.launch()
        run.ctx.sync()
        full += run.fetch(copy=True)
        run.close()
    nt = (n + tile - 1) // tile
    assert seen == {(i, j) for i in range(nt) for j in range(i, nt)}
    # every entry is written by exactly one part (the others left zeros): the sum is the matrix
    assert rel(full, want) <= 1e-9
    assert max(loads) <= 1.25 * (sum(loads) / world)          # round-robin on this forest: 1.6


# ---------------------------------------------------------------------------------------------------
# kernel variant of round 2: tree parts left in global memory (trees beyond the shared-memory staging limit)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,states", [(64, 0), (32, 0), (64, 3)])
def test_tree_parts_left_in_global_memory(precision, states):
    """The variant for trees beyond the shared-memory staging limit, forced on ordinary trees: identical results."""
    nwk = [synth.kingman_newick(90 + 20 * i, 9500 + i, states=states) for i in range(5)]
    flats = [FlatTree.from_newick(s, 3, "mean", states) for s in nwk]
    pk = PhyloKernel(precision=precision, n_states=states, **KAM)
    ctx = get_context()
    staged = pk.kernel_batch(flats)
    for option, path in ((1, "hbm-unstaged"), (2, "hbm-colglobal")):
        if option == 2 and states:
            continue                      # the column-global variant is for unlabelled trees
        try:
            ctx.set_option("unstaged", option)
            unstaged = pk.kernel_batch(flats)
            assert ctx.last_stats()["path"] == path
            g = pk.gram(flats, normalised=False)
        finally:
            ctx.set_option("unstaged", 0)
        if option == 1 and precision == 32:
            assert np.array_equal(unstaged, staged)      # same rows in the same order, one fp32 sum per rectangle in both
        # the staged fp64 kernels walk the rows with pk_rect2 (one running sum per thread instead of one per rectangle);
        # other CTA shapes (column-global: threads take several columns in turn) sum in another order
        assert rel(unstaged, staged) <= (1e-12 if precision == 64 else 1e-6)
        assert rel(g, staged) <= (1e-12 if precision == 64 else 1e-6)      # the Gram form evaluates (i, j) once and mirrors it


@pytest.mark.parametrize("precision", [64, 32])
def test_5000_tip_pair_beyond_the_staging_limit(precision):
    """Round 1 rejected trees whose packed parts exceed the shared memory of an SM (~3 500 tips); the reference has no size
    limit (phyloK2.py:158-209).  5 000 tips against 5 000 and against 3 000, checked against the C oracle."""
    a, b, c = synth.kingman_newick(5000, 9700), synth.yule_newick(5000, 9701), synth.kingman_newick(3000, 9702)
    pk = PhyloKernel(precision=precision, **KAM)
    fa, fb, fc = FlatTree.from_newick(a), FlatTree.from_newick(b), FlatTree.from_newick(c)
    got = pk.kernel_batch([fa, fb, fc], pairs=[(0, 1), (0, 0), (2, 1)])
    assert get_context().last_stats()["path"] == "hbm-unstaged"
    oa, ob, oc = oracle_flat(a), oracle_flat(b), oracle_flat(c)
    want = [c_oracle.kernel(oa, ob, 0.2, 2.0), c_oracle.kernel(oa, oa, 0.2, 2.0), c_oracle.kernel(oc, ob, 0.2, 2.0)]
    assert rel(got, want) <= TOL[precision]
    res = pk.score_batch(a, [b, c])
    assert abs(res["knorm"][0] - want[0] / math.sqrt(want[1] * c_oracle.kernel(ob, ob, 0.2, 2.0))) <= 10 * TOL[precision]


def test_repeated_uploads_of_a_tree_set_come_out_of_the_cached_image():
    """A set of >= 64 trees uploaded a second time in a row is kept as a page-locked packed image; later uploads are one
    copy out of it.  Same matrix every time, also when two sets alternate and when only part of the set is uploaded."""
    from kamphir_b200.gram import GramRunner
    sets = [synth.tree_set(70, 45, 9800, "mixed"), synth.tree_set(66, 60, 9900, "mixed")]
    flats = [[FlatTree.from_newick(s) for s in nwk] for nwk in sets]
    pk = PhyloKernel(**KAM)
    want = [pk.gram(f, normalised=True) for f in flats]
    runners = [GramRunner(f, 0.2, 2.0, 1.0, 64, True, 8) for f in flats]
    for order in ((0, 0, 0, 0), (1, 1, 1), (0, 1, 0, 1), (1, 1, 0, 0, 0)):
        for k in order:
            r = runners[k]
            r.upload()
            r.zero()
            r.launch()
            r.finish()
            assert np.array_equal(r.fetch(copy=True), want[k])
    # a part upload (what a rank of a multi-GPU run does) next to the plain one of the same trees
    ctx = get_context()
    for rep in range(3):
        part = Forest(ctx, flats[0], 64, part=(10, 40))
        full = Forest(ctx, flats[0], 64)
        lo, hi = 10, 40
        pa, na = part.device_range(lo, hi)
        pb, nb = full.device_range(lo, hi)
        assert na == nb and na > 0
        a, b = np.empty(na, np.uint8), np.empty(nb, np.uint8)
        from kamphir_b200 import _lib
        from kamphir_b200._lib import check
        check(_lib.lib().pk_memcpy_d2h(ctx._h, a.ctypes.data, pa, na))
        check(_lib.lib().pk_memcpy_d2h(ctx._h, b.ctypes.data, pb, nb))
        assert np.array_equal(a, b)
        part.close()
        full.close()
    for r in runners:
        r.close()


def test_kcoal_placement_small_and_large_replicates():
    """pk_score_newick_batch computes kcoal inside the flattening pass for small replicates and behind the launch for large
    ones (24 KB of text per tree decides): both give the numbers of the object path; a replicate with more internal nodes
    than the target is an error on either route (IndexError at kamphir.py:264 in the reference)."""
    from kamphir_b200 import FlatTree, PhyloKernel, synth
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0)
    for tips, nrep in ((60, 12), (900, 6)):
        obs = synth.kingman_newick(tips, 11)
        sims = [synth.kingman_newick(tips, 500 + i) for i in range(nrep)]
        assert (max(len(s) for s in sims) < 24576) == (tips == 60)
        a = pk.score_batch(obs, [x.encode() for x in sims], use_kcoal=True)
        fl = [FlatTree.from_newick(x) for x in sims]
        fo = FlatTree.from_newick(obs)
        want = np.array([t.kcoal(fo) for t in fl])
        assert np.array_equal(a["kcoal"], want)
        assert np.allclose(a["score"], a["knorm"] * want, rtol=1e-14, atol=0)
        bigger = synth.kingman_newick(tips + 5, 999)
        with pytest.raises(ValueError, match="more internal nodes"):
            pk.score_batch(obs, [x.encode() for x in sims[:3]] + [bigger.encode()], use_kcoal=True)
        pk.score_batch(obs, [x.encode() for x in sims[:3]] + [bigger.encode()])      # without kcoal it is a valid replicate
