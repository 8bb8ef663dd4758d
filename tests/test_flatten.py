"""C++ flattener (pk_tree_from_newick & co, host only) against the oracle's annotated trees: exact equality.

The oracle side is the Bio.Phylo stand-in + the restated prep (kamphir.py:279-282, phyloK2.py:101-146).  No GPU.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PREP, case_newick
from kamphir_b200 import FlatTree, PhyloKernel, synth
from oracle import phylok_oracle as po


def oracle_arrays(nwk, prep):
    f = po.prepare(po.read_newick(nwk), prep)
    return (np.array(f.prod, np.uint8), np.array(f.cprod, np.uint8), np.array(f.cidx, np.int32),
            np.array(f.bl, np.float64))


def assert_same(nwk, prep):
    flags, norm = PREP[prep]
    ft = FlatTree.from_newick(nwk, flags, norm)
    prod, cprod, cidx, bl = ft.export()
    o_prod, o_cprod, o_cidx, o_bl = oracle_arrays(nwk, prep)
    assert np.array_equal(prod, o_prod)
    assert np.array_equal(cprod, o_cprod)
    assert np.array_equal(cidx, o_cidx)
    assert np.array_equal(bl, o_bl)          # bit-exact, including the mean/median rescale
    return ft


def test_kat_trees_flatten_exactly(kat):
    seen = set()
    for c in kat["cases"]:
        for side in ("a", "b"):
            nwk = case_newick(c[side])
            key = (hash(nwk), c["prep"])
            if key in seen:
                continue
            seen.add(key)
            ft = assert_same(nwk, c["prep"])
            assert ft.n_internal == c["n_internal"][0 if side == "a" else 1]


@pytest.mark.parametrize("prep", ["raw", "kamphir:mean", "kamphir:median", "ladder:none"])
def test_synthetic_trees_flatten_exactly(prep):
    for nwk in (synth.tree_set(6, 37, 100, "mixed") + [synth.caterpillar_newick(30, 1), synth.balanced_newick(5, 2),
                                                       "(a:1,b:2);", "((a:1,b:2):0.5,c:3);"]):
        assert_same(nwk, prep)


def test_ladderize_is_stable_on_ties():
    # both children of the root have 2 tips: input order must be kept (Biopython: list.sort is stable)
    nwk = "((c:3,d:4):1,(a:1,b:2):2);"
    ft = FlatTree.from_newick(nwk, 2, "none")
    _, _, _, bl = ft.export()
    assert bl[0].tolist() == [3.0, 4.0] and bl[1].tolist() == [1.0, 2.0] and bl[2].tolist() == [1.0, 2.0]
    # and a real reordering: the single tip moves in front of the cherry
    ft = FlatTree.from_newick("((a:1,b:2):5,c:3);", 2, "none")
    _, cprod, cidx, bl = ft.export()
    assert bl[1].tolist() == [3.0, 5.0] and cidx[1].tolist() == [-1, 0] and cprod[1].tolist() == [0, 3]


def test_node_heights_and_kcoal_match_oracle():
    a, b = synth.kingman_newick(40, 1), synth.yule_newick(40, 2)
    fa, fb = FlatTree.from_newick(a), FlatTree.from_newick(b)
    ha = po.node_heights(po.read_newick(a))
    assert abs(fa.height - ha[0]) < 1e-15 * max(1, ha[0])
    np.testing.assert_allclose(fa.node_heights(), ha[1], rtol=1e-14, atol=1e-15)
    kc = po.kcoal(po.node_heights(po.read_newick(b))[1], ha[1])
    assert abs(fb.kcoal(fa) - kc) < 1e-14


def test_pair_counts_match_oracle():
    a, b = synth.kingman_newick(60, 7), synth.yule_newick(45, 8)
    fa, fb = FlatTree.from_newick(a), FlatTree.from_newick(b)
    assert fa.pair_counts(fb) == po.pair_counts(po.prep_kamphir(po.read_newick(a)), po.prep_kamphir(po.read_newick(b)))


def test_newick_dialect():
    base = FlatTree.from_newick("((A:1,B:2):3,C:4);", 0, "none").export()
    for text in ["(('A x':1,B:2)lbl:3,C:4)root;", "((A:1,B:2)[&c=1]:3,C:4.0E0);", " ( ( A : 1 , B :2):3,\nC:4 ) ;\n",
                 "((A:1.0,B:+2):0.3e1,C:4)99;", "(A:1,B:2):3,C:4;"]:
        got = FlatTree.from_newick(text, 0, "none").export()
        for x, y in zip(base, got):
            assert np.array_equal(x, y), text


def test_multi_tree_text_matches_single(kat):
    text = open(os.path.join(GOLDEN, "trees", "test.tree")).read()
    many = FlatTree.many_from_newick(text, 3, "mean", 0, 2)
    assert len(many) == 2
    for ft, nwk in zip(many, po._split_trees_text(text)):
        one = FlatTree.from_newick(nwk, 3, "mean")
        for x, y in zip(ft.export(), one.export()):
            assert np.array_equal(x, y)
    big = "\n".join(synth.tree_set(64, 20, 5))
    assert len(FlatTree.many_from_newick(big, 3, "mean", 0, 4)) == 64


@pytest.mark.parametrize("text,exc", [
    ("((a:1,b:2,c:3):1,d:1);", ValueError),        # polytomy          (reference: IndexError / partial)
    ("((a:1):1,d:1);", ValueError),                # unary node        (reference: IndexError phyloK2.py:188)
    ("a:1;", ValueError),                          # single tip        (reference: IndexError phyloK2.py:169)
    ("((a:1,b:2):1,d);", ValueError),              # missing length    (reference: TypeError phyloK2.py:146)
    ("((a:1,b:2):1,d:1;", ValueError),             # unbalanced parentheses
    ("((a:1,b:x):1,d:1);", ValueError),            # bad number
    ("((a:0,b:0):0,d:0);", ValueError),            # mean normalisation of an all-zero tree: ZeroDivisionError
    ("", ValueError),
])
def test_bad_trees_raise(text, exc):
    with pytest.raises(exc):
        FlatTree.from_newick(text, 3, "mean")


def test_tree_objects_and_python_prep_match_reference_semantics():
    """PhyloKernel.normalize_tree / annotate_tree on Bio.Phylo-like objects == oracle; flatten(object) == flatten(text)."""
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0)
    nwk = synth.kingman_newick(25, 9)
    for mode in ("mean", "median"):
        t = po.read_newick(nwk)
        t.root.branch_length = 0.
        t.ladderize()
        pk.normalize_tree(t, mode)
        pk.annotate_tree(t)
        o = po.prep_kamphir(po.read_newick(nwk), mode)
        nodes = t.get_nonterminals(order="postorder")
        assert [n.production for n in nodes] == o.prod and [n.bl for n in nodes] == o.bl
        assert [n.index for n in nodes] == list(range(len(nodes))) and [n.sqbl for n in nodes] == o.sqbl
        ft = pk.flatten(t)
        ref = FlatTree.from_newick(nwk, 3, mode)
        for x, y in zip(ft.export(), ref.export()):
            assert np.array_equal(x, y)


def test_labelled_classes():
    nwk = synth.kingman_newick(30, 4, states=2)
    ft = FlatTree.from_newick(nwk, 3, "mean", 2)
    assert ft.n_classes == 1 + 4 + 4 and ft.n_states == 2
    with pytest.raises(ValueError):
        FlatTree.from_newick(synth.kingman_newick(10, 4), 3, "mean", 2)      # labels without _<state>


def test_flat_tree_pickles():
    import pickle
    ft = FlatTree.from_newick(synth.kingman_newick(12, 3))
    for clone in (pickle.loads(pickle.dumps(ft)), pickle.loads(pickle.dumps(FlatTree.from_arrays(*ft.export()[2:])))):
        for x, y in zip(ft.export(), clone.export()):
            assert np.array_equal(x, y)
    pk = pickle.loads(pickle.dumps(PhyloKernel(decayFactor=0.3, precision=32)))
    assert pk.decayFactor == 0.3 and pk.precision == 32


def test_list_api_matches_single_calls():
    nwk = synth.tree_set(40, 25, 77) + ["  " + synth.kingman_newick(9, 3) + "  \n"]
    many = FlatTree.many_from_newick_list(nwk, 3, "mean", 0, 3)
    assert len(many) == len(nwk)
    for ft, s in zip(many, nwk):
        one = FlatTree.from_newick(s, 3, "mean")
        for x, y in zip(ft.export(), one.export()):
            assert np.array_equal(x, y)
    with pytest.raises(ValueError):
        FlatTree.many_from_newick_list(nwk[:3] + ["((a:1,b:2,c:3):1,d:1);"], 3, "mean")


def _child_flatten(seed):
    """Runs in a forked child: the parent's pool threads do not exist here, the library must build its own."""
    nwk = [synth.kingman_newick(60, seed * 100 + i) for i in range(40)]
    flats = FlatTree.many_from_newick("\n".join(nwk), nthreads=4)
    return [f.n_internal for f in flats]


@pytest.mark.timeout(120)
def test_worker_pool_survives_fork():
    """kamphir.py:638 creates an mp.Pool at module scope and forks workers that call the kernel object: the flattener's
    parked worker threads must not be assumed to exist in a forked child."""
    import multiprocessing as mp
    nwk = [synth.kingman_newick(60, i) for i in range(64)]
    parent = FlatTree.many_from_newick("\n".join(nwk), nthreads=4)        # pool threads now exist in this process
    assert len(parent) == 64
    with mp.get_context("fork").Pool(2) as pool:
        out = pool.map(_child_flatten, [1, 2, 3])
    assert all(x == [59] * 40 for x in out)
    again = FlatTree.many_from_newick("\n".join(nwk), nthreads=4)         # and the parent's pool still works
    assert [f.n_internal for f in again] == [59] * 64


def test_branch_length_text_parses_like_float():
    """The one-pass branch-length parser (8 digits per step, exact power-of-ten scaling, fallback for everything else)
    returns the double Python's float() -- i.e. Bio.Phylo's Newick reader -- returns, on awkward digit counts too."""
    import random
    rnd = random.Random(5)
    toks = []
    while len(toks) < 3000:
        k = rnd.choice([0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 15, 16, 17, 18, 19, 20, 24])
        ip = rnd.choice(['0', '7', '12', '123456789', '0000', '', '00012'])
        fr = ''.join(rnd.choice('0123456789') for _ in range(k))
        if rnd.random() < 0.3:
            fr = '0' * rnd.randint(1, 12) + fr
        ex = rnd.choice(['', 'e-05', 'E+3', 'e10', 'e-300', 'e0'])
        if not ip and not fr:
            continue
        tok = (ip + '.' + fr if fr or rnd.random() < 0.5 else (ip or '1')) + ex
        try:
            float(tok)
        except ValueError:
            continue
        if tok[0] in '.eE' and (len(tok) == 1 or tok[1] in 'eE'):
            continue
        toks.append(tok)
    for a, b in zip(toks[0::2], toks[1::2]):
        ft = FlatTree.from_newick("(x:%s,y:%s);" % (a, b), 0, "none")
        got = np.asarray(ft.export()[-1]).ravel()
        assert got[0] == float(a) and got[1] == float(b), (a, b, got)


def test_trees_from_multi_tree_text_pickle_completely():
    """many_from_newick trees keep no text of their own: the pickled form carries heights, height and scale, so kcoal and
    node_heights still work in a Pool worker (kamphir.py:27-32 pickles the whole object)."""
    import pickle
    nwk = synth.tree_set(5, 20, 31)
    target = FlatTree.from_newick(synth.kingman_newick(20, 99))
    for ft in FlatTree.many_from_newick("\n".join(nwk), 3, "mean"):
        clone = pickle.loads(pickle.dumps(ft))
        for x, y in zip(ft.export(), clone.export()):
            assert np.array_equal(x, y)
        assert np.array_equal(ft.node_heights(), clone.node_heights())
        assert (clone.height, clone.scale, clone.n_tips) == (ft.height, ft.scale, ft.n_tips)
        assert clone.kcoal(target) == ft.kcoal(target)


def test_kcoal_zero_norm_is_an_error_like_the_reference():
    # all node heights zero on one side: ZeroDivisionError at kamphir.py:269, ValueError here (never a silent NaN)
    z = FlatTree.from_newick("((a:0,b:0):0,c:0);", 3, "none")
    t = FlatTree.from_newick("((a:1,b:1):1,c:2);", 3, "none")
    with pytest.raises(ValueError):
        z.kcoal(t)
    with pytest.raises(ValueError):
        t.kcoal(z)


def test_flattened_form_is_cached_on_the_tree_object_until_it_is_prepared_again():
    """kamphir.py:290 passes the same observed tree to kernel() 2 * nreps times per step: it is walked once.  The cache
    follows the reference's snapshot semantics (kernel() reads node.bl as annotate_tree left it, phyloK2.py:169-185):
    normalize_tree / annotate_tree drop it."""
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0)
    t = po.read_newick(synth.kingman_newick(15, 5))
    t.root.branch_length = 0.
    t.ladderize()
    pk.normalize_tree(t, "mean")
    pk.annotate_tree(t)
    f1 = pk.flatten(t)
    assert pk.flatten(t) is f1
    pk.normalize_tree(t, "median")            # rescales the live branch lengths: the snapshot is stale until annotate
    assert getattr(t, "_pk_flat", None) is None
    pk.annotate_tree(t)
    f2 = pk.flatten(t)
    assert f2 is not f1 and pk.flatten(t) is f2
    assert not np.array_equal(f1.export()[3], f2.export()[3])
    nodes = t.get_nonterminals(order="postorder")
    assert f2.export()[3].tolist() == [n.bl for n in nodes]


def test_kernel_matrix_consumers(tmp_path):
    """normalised_matrix / save_matrix / kernel_pca work on a computed matrix (here: filled by the C oracle, no GPU):
    the component coordinates reproduce the double-centred matrix and are ordered by variance."""
    from oracle import c_oracle
    nwk = synth.tree_set(12, 20, 77)
    flats = [c_oracle.CFlat(po.prep_kamphir(po.read_newick(s))) for s in nwk]
    pk = PhyloKernel(decayFactor=0.2, gaussFactor=2.0)
    with pytest.raises(ValueError):
        pk.normalised_matrix()
    pk.kmat = np.array([[c_oracle.kernel(a, b, 0.2, 2.0) for b in flats] for a in flats])
    pk.is_kmat_computed = True
    kn = pk.normalised_matrix()
    assert np.allclose(np.diag(kn), 1.0) and np.allclose(kn, kn.T) and kn.max() <= 1.0 + 1e-12
    pk.save_matrix(tmp_path / "k.npy")
    pk.save_matrix(tmp_path / "kn.npy", normalised=True)
    assert np.array_equal(np.load(tmp_path / "k.npy"), pk.kmat) and np.array_equal(np.load(tmp_path / "kn.npy"), kn)
    coords, frac = pk.kernel_pca(n_components=11)
    j = np.eye(12) - 1.0 / 12
    assert np.allclose(coords @ coords.T, j @ kn @ j, atol=1e-10)          # 11 components span the centred matrix
    assert np.all(np.diff(frac) <= 1e-12) and abs(frac.sum() - 1.0) < 1e-9
    assert np.allclose(coords.sum(axis=0), 0.0, atol=1e-9)
    c2, f2 = pk.kernel_pca(2)
    assert np.allclose(c2, coords[:, :2]) and np.allclose(f2, frac[:2])
    with pytest.raises(ValueError):
        pk.kernel_pca(13)
