"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200:  pytest -m gpu

Tolerances (BASELINE.json north_star): relative 1e-9 in fp64 mode, 1e-5 in fp32 mode, on K and on normalised K.
The reference values are either the committed goldens (produced by the reference's own phyloK2.py) or the C oracle
on the same seeded inputs.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from conftest import PREP, case_newick
from kamphir_b200 import FlatTree, PhyloKernel, get_context, synth
from oracle import c_oracle
from oracle import phylok_oracle as po

pytestmark = pytest.mark.gpu
TOL = {64: 1e-9, 32: 1e-5}
KAM = dict(decayFactor=0.2, gaussFactor=2.0)


def oracle_flat(nwk, prep="kamphir:mean"):
    return c_oracle.CFlat(po.prepare(po.read_newick(nwk), prep))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(autouse=True)
def _default_config():
    get_context().configure(0, 0, 0)
    yield
    get_context().configure(0, 0, 0)


@pytest.mark.parametrize("precision", [64, 32])
def test_golden_vectors(kat, precision):
    """Every KAT of tests/golden/kernel_kat.json (values from the reference itself)."""
    worst = 0.0
    for c in kat["cases"]:
        flags, norm = PREP[c["prep"]]
        pk = PhyloKernel(decayFactor=c["lam"], gaussFactor=c["tau"], sigma=c["sigma"], precision=precision)
        a = FlatTree.from_newick(case_newick(c["a"]), flags, norm)
        b = FlatTree.from_newick(case_newick(c["b"]), flags, norm)
        k = pk.kernel(a, b)
        err = abs(k - c["K"]) / abs(c["K"])
        worst = max(worst, err)
        assert err <= TOL[precision], (c["name"], k, c["K"], err)
    print("worst relative error fp%d: %.3g" % (precision, worst))


def test_reference_unittests_through_tree_objects():
    """tests/unittests.py:83-101 as written there: Bio.Phylo-like objects, no prep, lambda 0.5, tau 1."""
    import math
    pk = PhyloKernel(decayFactor=0.5, gaussFactor=1)
    t1 = po.read_newick("( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;")
    t2 = po.read_newick("( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;")
    assert abs(pk.kernel(t1, t2) - 1.125 * (1 + math.exp(-0.0625))) < 1e-14
    assert hasattr(t1.root, "production") and t1.root.index == 2        # annotated lazily like phyloK2.py:169-172
    # kamphir-style use: prep on the object with the class's own methods, then kernel(tree, tree)
    pk2 = PhyloKernel(**KAM)
    nwk = synth.kingman_newick(60, 17)
    t = po.read_newick(nwk)
    t.root.branch_length = 0.
    t.ladderize()
    pk2.normalize_tree(t, "mean")
    pk2.annotate_tree(t)
    f = oracle_flat(nwk)
    assert rel(pk2.kernel(t, t), c_oracle.kernel(f, f, 0.2, 2.0)) < 1e-12
    # striped mode: ranks sum to the full kernel (the reference's own version is wrong, phyloK2.py:223)
    assert sum(pk2.kernel(t, t, r, 4) for r in range(4)) == pk2.kernel(t, t)


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("force_hbm", [2, 1, 0])
def test_random_pairs_against_c_oracle(precision, force_hbm):
    """Ragged sizes, both storage paths (C in shared memory / C in HBM slabs)."""
    get_context().configure(0, 0, force_hbm)
    rng = np.random.default_rng(5)
    sizes = [2, 3, 4, 5, 7, 16, 33, 64, 65, 100, 129, 200, 201, 220]
    nwk = [synth.kingman_newick(n, 100 + i) if i % 2 else synth.yule_newick(n, 100 + i) for i, n in enumerate(sizes)]
    nwk += [synth.caterpillar_newick(50, 3), synth.balanced_newick(6, 4)]
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, len(nwk), size=(60, 2))] + [(i, i) for i in range(len(nwk))]
    pk = PhyloKernel(precision=precision, **KAM)
    got = pk.kernel_batch(flats, flats, pairs=pairs)
    want = [c_oracle.kernel(orac[a], orac[b], 0.2, 2.0) for a, b in pairs]
    assert rel(got, want) <= TOL[precision]
    stats = get_context().last_stats()
    assert stats["launches"] >= 1
    if force_hbm:
        assert stats["path"] == ("hbm" if force_hbm == 1 else "smem")


@pytest.mark.parametrize("threads", [32, 64, 96, 128, 192, 224, 256, 320, 352, 384, 512, 576, 768, 1024])   # every kernel tier
def test_every_cta_shape_gives_the_same_answer(threads):
    """The CTA size selects the kernel instantiation (register budget, rows in flight, resident CTAs); a pair's cells do
    not depend on it, only the order of the final sum does."""
    for precision, states in [(64, 0), (32, 0), (64, 2)]:
        nwk = synth.tree_set(6, 120, 900, "mixed", states, [0.5, 0.5] if states else None)
        flats = [FlatTree.from_newick(s, 3, "mean", states) for s in nwk]
        pk = PhyloKernel(precision=precision, n_states=states, **KAM)
        get_context().configure(0, 0, 0)
        base = pk.kernel_batch(flats)
        get_context().configure(threads, 0, 0)
        a = pk.kernel_batch(flats)
        get_context().configure(threads, 2, 2 if threads <= 256 else 1)
        b = pk.kernel_batch(flats)
        assert rel(a, base) < 1e-13 and rel(b, base) < 1e-13, (precision, states)


@pytest.mark.parametrize("precision,states", [(32, 0), (64, 3), (32, 3), (64, 0)])
def test_small_pairs_after_large_pairs_on_the_same_cta(precision, states):
    """One resident CTA per SM with the C cells in HBM slabs works through a list that alternates large and small trees:
    the staged tree of a small pair is followed in shared memory by what the large one left behind, and the unclamped
    prefetch of the fp32 / labelled kernel variants (PK_OVERRUN) reads the records after a rectangle -- other rows of the
    same tree, then the never-matching pad records -- and may address the slab's slack.  Results must not depend on it."""
    sizes = [300, 7, 150, 3, 260, 12, 2, 90, 280, 5, 33, 310]
    nwk = [synth.kingman_newick(n, 700 + i, states=states) if i % 3 else synth.yule_newick(n, 700 + i, states=states)
           for i, n in enumerate(sizes)]
    flats = [FlatTree.from_newick(s, 3, "mean", states) for s in nwk]
    pk = PhyloKernel(precision=precision, n_states=states, **KAM)
    pairs = [(i, j) for i in range(len(sizes)) for j in range(len(sizes))] * 3      # 432 work items > resident CTAs
    get_context().configure(0, 1, 1)
    got = pk.kernel_batch(flats, flats, pairs)
    if states:
        of = [po.prep_kamphir(po.read_newick(s), "mean", po.state_from_name) for s in nwk]
        want = {}
        for i, j in set(pairs):
            if i <= j:
                want[(i, j)] = want[(j, i)] = po.kernel_arrays_labelled(of[i], of[j], 0.2, 2.0, 1.0)
        want = np.array([want[p] for p in pairs])
    else:
        orac = [oracle_flat(s) for s in nwk]
        want = np.array([c_oracle.kernel(orac[i], orac[j], 0.2, 2.0, 1.0) for i, j in pairs])
    assert rel(got, want) < TOL[precision]
    n = len(sizes) ** 2
    assert np.array_equal(got[:n], got[n:2 * n]) and np.array_equal(got[:n], got[2 * n:])      # deterministic


def test_parameters_sigma_lambda_tau():
    nwk = synth.tree_set(4, 80, 40)
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    for lam, tau, sig in [(0.1, 1.0, 1.0), (0.5, 1.0, 1.0), (0.3, 0.7, 0.25), (0.2, 50.0, 2.0), (0.0, 1.0, 1.0)]:
        pk = PhyloKernel(decayFactor=lam, gaussFactor=tau, sigma=sig)
        got = pk.kernel_batch(flats)
        want = np.array([[c_oracle.kernel(a, b, lam, tau, sig) for b in orac] for a in orac])
        if lam == 0.0:
            assert np.all(got == 0.0) and np.all(want == 0.0)
        else:
            assert rel(got, want) < 1e-9


@pytest.mark.parametrize("precision", [64, 32])
def test_500_tip_pairs(precision):
    """BASELINE.json's headline size (configs[2], configs[3])."""
    nwk = [synth.kingman_newick(500, 3000), synth.yule_newick(500, 3001), synth.kingman_newick(500, 3002)]
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    pk = PhyloKernel(precision=precision, **KAM)
    got = pk.kernel_batch(flats)
    want = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in orac] for a in orac])
    assert rel(got, want) <= TOL[precision]
    assert get_context().last_stats()["path"] == "hbm"
    assert rel(got, got.T) < (1e-12 if precision == 64 else 1e-6)     # K(a,b) == K(b,a) up to summation order


@pytest.mark.parametrize("precision", [64, 32])
def test_2000_tip_pairs_pangea_size(precision):
    """configs[4]: PANGEA-scale trees, C matrix resident in HBM."""
    nwk = [synth.kingman_newick(2000, 5000), synth.kingman_newick(2000, 5001), synth.yule_newick(1500, 5002)]
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    pk = PhyloKernel(precision=precision, **KAM)
    pairs = [(0, 1), (0, 0), (1, 2), (2, 0)]
    got = pk.kernel_batch(flats, flats, pairs=pairs)
    want = [c_oracle.kernel(orac[a], orac[b], 0.2, 2.0) for a, b in pairs]
    assert rel(got, want) <= TOL[precision]


def test_deep_and_degenerate_trees():
    """Caterpillars (one wavefront level per node), a lone cherry, very different sizes in one launch."""
    nwk = [synth.caterpillar_newick(400, 1), synth.caterpillar_newick(399, 2), "(a:0.3,b:0.7);",
           synth.balanced_newick(8, 3), synth.kingman_newick(700, 9)]
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    assert flats[0].n_levels == 399 and flats[2].n_internal == 1
    pk = PhyloKernel(**KAM)
    got = pk.kernel_batch(flats)
    want = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in orac] for a in orac])
    assert rel(got, want) < 1e-9


def test_unladderized_trees_tip_in_either_slot():
    """kernel() does not re-prepare: production-2 nodes with the tip in slot 1 must not match slot-0 tips."""
    nwk = synth.tree_set(6, 90, 77)
    flats = [FlatTree.from_newick(s, 0, "none") for s in nwk]
    orac = [oracle_flat(s, "raw") for s in nwk]
    pk = PhyloKernel(decayFactor=0.35, gaussFactor=1.5)
    got = pk.kernel_batch(flats)
    want = np.array([[c_oracle.kernel(a, b, 0.35, 1.5) for b in orac] for a in orac])
    assert rel(got, want) < 1e-9


@pytest.mark.parametrize("precision", [64, 32])
def test_wide_dynamic_range_of_the_gaussian(precision):
    """Un-normalised trees on very different length scales: exponents from 0 to far beyond exp(-700) in one launch
    (fp64: the table-based Gaussian clamps its exponent instead of underflowing; the reference's math.exp gives 0)."""
    import re
    base = synth.tree_set(4, 120, 8100, "mixed")

    def scaled(nwk, f):
        return re.sub(r":([0-9.eE+-]+)", lambda m: ":%.10g" % (float(m.group(1)) * f), nwk)

    nwk = [scaled(base[0], 1.0), scaled(base[1], 40.0), scaled(base[2], 0.01), scaled(base[3], 3000.0), scaled(base[0], 40.0)]
    flats = [FlatTree.from_newick(s, 3, "none") for s in nwk]
    orac = [oracle_flat(s, "kamphir:none") for s in nwk]
    for tau in (2.0, 0.05):
        pk = PhyloKernel(decayFactor=0.2, gaussFactor=tau, normalize="none", precision=precision)
        got = pk.kernel_batch(flats)
        want = np.array([[c_oracle.kernel(a, b, 0.2, tau) for b in orac] for a in orac])
        assert np.all(np.isfinite(got)) and np.all(got >= 0)
        # relative where the reference is representable, absolute (far below any meaningful K) where it underflows
        assert np.all(np.abs(got - want) <= TOL[precision] * want + 1e-290)


def test_parameter_domain():
    """gaussFactor must be positive and decayFactor non-negative; decayFactor = 0 gives K = 0 exactly."""
    a = "((a:1,b:2):1,(c:1,d:3):2);"
    for kw in (dict(gaussFactor=-1.0), dict(decayFactor=-0.1), dict(gaussFactor=float("nan"))):
        with pytest.raises(ValueError):
            PhyloKernel(**kw).kernel(a, a)
    assert PhyloKernel(decayFactor=0.0).kernel(a, a) == 0.0
    assert PhyloKernel(decayFactor=0.0, precision=32).kernel(a, a) == 0.0


def test_score_batch_matches_kamphir_compute():
    """configs[0]/[1]: one ABC proposal = nreps simulated trees against the observed tree (kamphir.py:249-306,434-450)."""
    obs = synth.kingman_newick(100, 1)
    sims = [synth.kingman_newick(100, 200 + i) for i in range(12)]
    mean, scores = po.kamphir_evaluate(sims, obs, decayFactor=0.2, gaussFactor=2.0, normalize="mean")
    pk = PhyloKernel(normalize="mean", **KAM)
    res = pk.score_batch(obs, sims, use_kcoal=True, kadj=1.0)
    assert rel(res["score"], scores) < 1e-9 and abs(res["mean_score"] - mean) < 1e-9 * mean
    fo = oracle_flat(obs)
    assert rel(res["ref_denom"], c_oracle.kernel(fo, fo, 0.2, 2.0)) < 1e-12
    # pre-computed ref_denom is passed through (kamphir.py:157 computes it once)
    res2 = pk.score_batch(obs, sims, ref_denom=res["ref_denom"])
    assert np.array_equal(res2["knorm"], res["knorm"])
    assert np.all(res["knorm"] > 0) and np.all(res["knorm"] <= 1 + 1e-12)


def test_score_batch_text_path_equals_object_path():
    """Newick text goes through pk_score_newick_batch (one call), flattened trees through pk_score_batch + pk_kcoal_batch:
    same numbers, also for labelled trees and the median normalisation; a bad replicate raises."""
    for S, norm in ((0, "mean"), (0, "median"), (2, "mean")):
        obs = synth.kingman_newick(70, 5, states=S)
        sims = [synth.kingman_newick(70, 300 + i, states=S) for i in range(9)]
        pk = PhyloKernel(normalize=norm, n_states=S, **KAM)
        a = pk.score_batch(obs, [x.encode() for x in sims], use_kcoal=True, kadj=0.7)
        fl = [FlatTree.from_newick(x, 3, norm, S) for x in sims]
        b = pk.score_batch(obs, fl, use_kcoal=True, kadj=0.7)
        for key in ("k", "denom", "knorm", "kcoal"):
            assert np.array_equal(a[key], b[key]), key
        assert rel(a["score"], b["score"]) < 1e-14           # pow() of the C library vs numpy's
        assert a["mean_score"] == pytest.approx(b["mean_score"], rel=1e-15) and a["ref_denom"] == b["ref_denom"]
        c = pk.score_batch(obs, sims)                      # without kcoal: the mean is the mean of knorm
        assert "score" not in c and c["mean_score"] == pytest.approx(float(np.mean(a["knorm"])), rel=1e-14)
    with pytest.raises(ValueError):
        PhyloKernel(**KAM).score_batch("(a:1,b:2);", ["(a:1,b:2);", "((a:1,b:2,c:3):1,d:1);"])


@pytest.mark.parametrize("precision", [64, 32])
def test_gram_matrix(precision):
    """configs[3] in miniature: symmetry, unit diagonal, entries against the oracle, raw == compute_matrix semantics."""
    nwk = synth.tree_set(40, 60, 4000, "mixed")
    flats = [FlatTree.from_newick(s) for s in nwk]
    orac = [oracle_flat(s) for s in nwk]
    pk = PhyloKernel(precision=precision, **KAM)
    raw = pk.gram(flats, normalised=False)
    want = np.array(po.gram([po.prep_kamphir(po.read_newick(s)) for s in nwk[:8]], 0.2, 2.0))
    assert rel(raw[:8, :8], want) <= TOL[precision]
    assert np.array_equal(raw, raw.T)
    nrm = pk.gram(flats, normalised=True)
    assert np.array_equal(nrm, nrm.T) and np.allclose(np.diag(nrm), 1.0, rtol=0, atol=1e-12)
    d = np.sqrt(np.diag(raw))
    assert rel(nrm, raw / np.outer(d, d)) <= 1e-12
    rng = np.random.default_rng(0)
    for i, j in rng.integers(0, 40, size=(20, 2)):
        assert abs(raw[i, j] - c_oracle.kernel(orac[i], orac[j], 0.2, 2.0)) <= TOL[precision] * raw[i, j]


def test_gram_at_headline_size_properties():
    """configs[3] shape at a size the oracle still samples quickly: 160 x 500-tip trees (12 880 DPs): symmetry, unit
    diagonal, normalised = raw / sqrt(diag diag), a Cauchy-Schwarz bound, and sampled entries against the C oracle."""
    nwk = synth.tree_set(160, 500, 3000, "mixed")
    flats = FlatTree.many_from_newick("\n".join(nwk))
    pk = PhyloKernel(**KAM)
    raw = pk.gram(flats, normalised=False)
    nrm = pk.gram(flats, normalised=True)
    assert np.array_equal(raw, raw.T) and np.array_equal(nrm, nrm.T)
    assert np.allclose(np.diag(nrm), 1.0, rtol=0, atol=1e-12)
    d = np.sqrt(np.diag(raw))
    assert rel(nrm, raw / np.outer(d, d)) <= 1e-12
    assert np.all(nrm > 0) and np.all(nrm <= 1 + 1e-12)          # K is a positive-definite kernel: |K_ij| <= sqrt(K_ii K_jj)
    assert get_context().last_stats()["path"] == "hbm"
    rng = np.random.default_rng(1)
    for i, j in rng.integers(0, 160, size=(12, 2)):
        a, b = oracle_flat(nwk[i]), oracle_flat(nwk[j])
        assert abs(raw[i, j] - c_oracle.kernel(a, b, 0.2, 2.0)) <= 1e-9 * raw[i, j]


def test_pangea_size_proposal():
    """configs[4] shape: 2000-tip replicates against a 2000-tip target through the one-call path; k and denom against the
    C oracle, the normalised score from them as kamphir.py:300 defines it, the mean as kamphir.py:450."""
    obs = synth.kingman_newick(2000, 2000)
    sims = [synth.kingman_newick(2000, 2100 + i) for i in range(6)]
    pk = PhyloKernel(normalize="mean", **KAM)
    res = pk.score_batch(obs, [s.encode() for s in sims], use_kcoal=True)
    fo = oracle_flat(obs)
    ref = c_oracle.kernel(fo, fo, 0.2, 2.0)
    assert rel(res["ref_denom"], ref) < 1e-12
    for s, k, dn, kh in zip(sims, res["k"], res["denom"], res["knorm"]):
        fs = oracle_flat(s)
        want_k, want_d = c_oracle.kernel(fo, fs, 0.2, 2.0), c_oracle.kernel(fs, fs, 0.2, 2.0)
        assert abs(k - want_k) <= 1e-9 * want_k and abs(dn - want_d) <= 1e-9 * want_d
        assert abs(kh - np.exp(np.log(want_k) - 0.5 * (np.log(ref) + np.log(want_d)))) <= 1e-9
    assert res["mean_score"] == pytest.approx(float(np.mean(res["knorm"] * res["kcoal"])), rel=1e-14)


def test_gram_parts_tile_the_matrix():
    """The sharded form: parts are disjoint, cover the upper triangle, and agree with the one-part result."""
    import ctypes
    from kamphir_b200 import Forest, _lib
    from kamphir_b200._lib import check, pk_params
    nwk = synth.tree_set(37, 50, 6000, "mixed")
    flats = [FlatTree.from_newick(s) for s in nwk]
    n = len(flats)
    ctx = get_context()
    forest = Forest(ctx, flats, 64)
    L = _lib.lib()
    p = pk_params(0.2, 2.0, 1.0, 64, 0)
    full = PhyloKernel(**KAM).gram(flats, normalised=True)
    d_out, d_diag = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.pk_device_alloc(ctx._h, 8 * n * n, ctypes.byref(d_out)))
    check(L.pk_device_alloc(ctx._h, 8 * n, ctypes.byref(d_diag)))
    total = 0
    acc = np.zeros((n, n))
    for part in range(3):
        check(L.pk_memset_d(ctx._h, d_out, 0, 8 * n * n))
        check(L.pk_forest_gram_part(ctx._h, forest._h, ctypes.byref(p), 1, 4, part, 3, d_out, n, d_diag))
        got = np.empty((n, n))
        check(L.pk_memcpy_d2h(ctx._h, got.ctypes.data, d_out, 8 * n * n))
        assert not np.any((acc != 0) & (got != 0))           # disjoint
        acc += got
        # the forest-aware deal (by cost): its tile list, its DP count and what the kernel wrote agree
        nt = ctypes.c_int64(0)
        check(L.pk_forest_gram_part_tiles(forest._h, 4, part, 3, None, ctypes.byref(nt)))
        tiles = np.zeros((nt.value, 2), np.int32)
        check(L.pk_forest_gram_part_tiles(forest._h, 4, part, 3, tiles.ctypes.data, ctypes.byref(nt)))
        dps = sum((min(n, 4 * ti + 4) - 4 * ti) * (min(n, 4 * tj + 4) - 4 * tj) if ti != tj
                  else (min(n, 4 * ti + 4) - 4 * ti) * (min(n, 4 * ti + 4) - 4 * ti + 1) // 2 for ti, tj in tiles)
        assert dps == forest.gram_stats(4, part, 3)["dps"] == int(np.count_nonzero(np.triu(got)))
        total += dps
    assert total == n * (n + 1) // 2 and np.array_equal(acc, full)
    check(L.pk_device_free(ctx._h, d_out))
    check(L.pk_device_free(ctx._h, d_diag))


def test_forest_uploaded_in_parts_equals_plain_upload():
    """The multi-GPU replication path on one GPU: two forests each pack half of the trees, the missing halves are copied
    across (what the NCCL broadcasts do between ranks), and both then give the Gram matrix of a plain upload."""
    import ctypes
    from kamphir_b200 import Forest, _lib
    from kamphir_b200._lib import check, pk_params
    nwk = synth.tree_set(21, 70, 8800, "mixed")
    flats = [FlatTree.from_newick(s) for s in nwk]
    n, cut = len(flats), 9
    ctx = get_context()
    L = _lib.lib()
    plain = Forest(ctx, flats, 64)
    fa = Forest(ctx, flats, 64, part=(0, cut))
    fb = Forest(ctx, flats, 64, part=(cut, n))
    assert fa.nbytes() == fb.nbytes() == plain.nbytes()
    for dst, src, lo, hi in ((fa, fb, cut, n), (fb, fa, 0, cut)):
        sp, sb = src.device_range(lo, hi)
        dp, db = dst.device_range(lo, hi)
        assert sb == db > 0
        host = np.empty(sb, np.uint8)
        check(L.pk_memcpy_d2h(ctx._h, host.ctypes.data, ctypes.c_void_p(sp), sb))
        check(L.pk_memcpy_h2d(ctx._h, ctypes.c_void_p(dp), host.ctypes.data, sb))
    p = pk_params(0.2, 2.0, 1.0, 64, 0)
    outs = []
    for f in (plain, fa, fb):
        d_out, d_diag = ctypes.c_void_p(), ctypes.c_void_p()
        check(L.pk_device_alloc(ctx._h, 8 * n * n, ctypes.byref(d_out)))
        check(L.pk_device_alloc(ctx._h, 8 * n, ctypes.byref(d_diag)))
        check(L.pk_forest_gram_part(ctx._h, f._h, ctypes.byref(p), 1, 4, 0, 1, d_out, n, d_diag))
        got = np.empty((n, n))
        check(L.pk_memcpy_d2h(ctx._h, got.ctypes.data, d_out, 8 * n * n))
        outs.append(got)
        check(L.pk_device_free(ctx._h, d_out))
        check(L.pk_device_free(ctx._h, d_diag))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert f.gram_stats(4) == plain.gram_stats(4)
    for f in (plain, fa, fb):
        f.close()


def test_labelled_tip_states_against_extension_oracle():
    """configs[2] extension (no reference implementation, parity unpinned): class = production + tip-child states."""
    S = 2
    nwk = [synth.kingman_newick(80, 50 + i, states=S, prevalence=[0.5, 0.5]) for i in range(5)]

    def classes(flat):
        out = []
        for i in range(flat.n):
            st = flat.state[i]
            tips = [s for s in (0, 1) if flat.cidx[i][s] < 0]
            if not tips:
                out.append(0)
            elif len(tips) == 1:
                out.append(1 + tips[0] * S + st[tips[0]])
            else:
                out.append(1 + 2 * S + st[0] * S + st[1])
        return out

    pk = PhyloKernel(n_states=S, **KAM)
    flats = [FlatTree.from_newick(s, 3, "mean", S) for s in nwk]
    got = pk.kernel_batch(flats)
    of = [po.prep_kamphir(po.read_newick(s), "mean", po.state_from_name) for s in nwk]
    want = np.array([[po.kernel_arrays_labelled(a, b, 0.2, 2.0) for b in of] for a in of])
    assert rel(got, want) < 1e-9
    # the C++ flattener's class ids equal the test's formula, and with one state the labelled kernel is the plain one
    for ft, f in zip(flats, of):
        assert (ft.export()[0] - 1).tolist() == classes(f)
    one = [synth.kingman_newick(80, 50 + i, states=1) for i in range(3)]
    lab = PhyloKernel(n_states=1, **KAM).kernel_batch([FlatTree.from_newick(s, 3, "mean", 1) for s in one])
    plain = PhyloKernel(**KAM).kernel_batch([FlatTree.from_newick(s) for s in one])
    assert rel(lab, plain) < 1e-13


@pytest.mark.parametrize("S", [3, 7])
def test_labelled_many_classes(S):
    """Up to 1 + 2S + S^2 = 64 production classes (S = 7): narrow class rectangles dealt to the warps by chunks, one
    descriptor per class and level; Gram form and pair form agree with the extension oracle."""
    nwk = [synth.kingman_newick(150, 900 + i, states=S) for i in range(3)] + [synth.yule_newick(260, 950, states=S)]
    flats = [FlatTree.from_newick(s, 3, "mean", S) for s in nwk]
    of = [po.prep_kamphir(po.read_newick(s), "mean", po.state_from_name) for s in nwk]
    want = np.array([[po.kernel_arrays_labelled(a, b, 0.2, 2.0) for b in of] for a in of])
    for precision in (64, 32):
        pk = PhyloKernel(n_states=S, precision=precision, **KAM)
        got = pk.kernel_batch(flats)
        assert rel(got, want) <= TOL[precision]
        raw = pk.gram(flats, normalised=False)
        assert rel(raw, want) <= TOL[precision]
        get_context().configure(32, 0, 0)          # a one-warp CTA: fewer threads than classes (S >= 5)
        assert rel(pk.kernel_batch(flats), want) <= TOL[precision]
        get_context().configure(0, 0, 0)


def test_load_trees_from_file_and_compute_matrix(tmp_path):
    """The reference's broken matrix surface (phyloK2.py:68-99,148-156) works here."""
    nwk = synth.tree_set(5, 30, 10)
    path = tmp_path / "trees.nwk"
    path.write_text("\n".join(nwk) + "\n")
    pk = PhyloKernel(**KAM)
    with open(path) as fh:
        pk.load_trees_from_file(fh)
    assert pk.ntrees == 5 and not pk.is_kmat_computed
    pk.compute_matrix()
    orac = [oracle_flat(s) for s in nwk]
    want = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in orac] for a in orac])
    assert pk.is_kmat_computed and rel(pk.kmat, want) < 1e-9
    d = np.sqrt(np.diag(want))
    assert rel(pk.normalised_matrix(), want / np.outer(d, d)) < 1e-9
    pk.save_matrix(tmp_path / "kmat.npy")
    assert np.array_equal(np.load(tmp_path / "kmat.npy"), pk.kmat)
    coords, frac = pk.kernel_pca(2)
    assert coords.shape == (5, 2) and frac[0] >= frac[1] >= 0.0


def test_errors_surface_as_python_exceptions():
    pk = PhyloKernel(**KAM)
    with pytest.raises(ValueError):
        pk.kernel("((a:1,b:2,c:3):1,d:1);", "(a:1,b:2);")
    with pytest.raises(ValueError):
        PhyloKernel(gaussFactor=0.0).kernel("(a:1,b:2);", "(a:1,b:2);")
    lab = FlatTree.from_newick(synth.kingman_newick(10, 1, states=2), 3, "mean", 2)
    with pytest.raises(ValueError):
        pk.kernel(lab, "(a:1,b:2);")          # different production-class schemes
    sims = [synth.kingman_newick(12, i) for i in range(3)]
    with pytest.raises(ValueError, match="math domain"):      # K = 0: math.log(0) at kamphir.py:300
        PhyloKernel(decayFactor=0.0).score_batch(sims[0], sims)
    with pytest.raises(ValueError, match="math domain"):
        PhyloKernel(decayFactor=0.0).score_batch(sims[0], [FlatTree.from_newick(s) for s in sims])
