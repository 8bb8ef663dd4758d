"""The oracle (oracle/phylok_oracle.py + oracle/phylok_oracle.c) against the reference's goldens.

Pins: (1) the literals of /root/reference/tests/unittests.py:93-113; (2) tests/golden/kernel_kat.json, whose values
were produced by the reference's unmodified phyloK2.py (tests/golden/make_golden.py).  CPU only.
"""
import math

import pytest

from conftest import case_newick
from oracle import c_oracle
from oracle import phylok_oracle as po

T1 = "( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;"
T2 = "( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;"


def test_reference_unittest_kernel_exact():
    # tests/unittests.py:93-95 (assertEqual, i.e. bit-exact)
    k = po.kernel(po.read_newick(T1), po.read_newick(T2), decayFactor=0.5, gaussFactor=1)
    assert k == 1.125 * (1 + math.exp(-0.0625))


def test_reference_unittest_kernel_large(kat):
    # tests/unittests.py:97-101 (assertAlmostEqual, 7 places)
    c = [c for c in kat["cases"] if c["name"] == "unittest_kernel_large"][0]
    k = po.kernel(po.read_newick(case_newick(c["a"])), po.read_newick(case_newick(c["b"])), 0.5, 1)
    assert round(k - 3042.58677467, 7) == 0


@pytest.mark.parametrize("mode", ["mean", "median"])
def test_reference_unittest_normalize(mode):
    # tests/unittests.py:103-113: every non-root branch of T1 divided by 0.375
    t = po.read_newick(T1)
    t.root.branch_length = 0.
    expected = [c.branch_length / 0.375 for c in t.find_clades(order="postorder")][:-1]
    po.normalize_tree(t, mode)
    assert [c.branch_length for c in t.find_clades(order="postorder")][:-1] == expected


def test_normalize_golden(kat):
    g = kat["normalize"]
    t = po.read_newick(g["newick"])
    t.root.branch_length = 0.
    po.normalize_tree(t, g["mode"])
    assert [c.branch_length for c in t.find_clades(order="postorder")][:-1] == g["postorder_non_root"]


def test_python_oracle_matches_reference_bitwise(kat):
    """Same loop order and arithmetic as phyloK2.py:158-209 -> identical floats."""
    for c in kat["cases"]:
        if max(c["n_internal"]) > 320:
            continue          # pure-Python loops: keep the CPU suite quick; the C oracle covers the big ones
        a = po.prepare(po.read_newick(case_newick(c["a"])), c["prep"])
        b = po.prepare(po.read_newick(case_newick(c["b"])), c["prep"])
        k = po.kernel_arrays(a, b, c["lam"], c["tau"], c["sigma"])
        assert k == c["K"], c["name"]


def test_c_oracle_matches_reference(kat):
    for c in kat["cases"]:
        a = po.prepare(po.read_newick(case_newick(c["a"])), c["prep"])
        b = po.prepare(po.read_newick(case_newick(c["b"])), c["prep"])
        k, (npairs, m, r) = c_oracle.kernel(c_oracle.CFlat(a), c_oracle.CFlat(b), c["lam"], c["tau"], c["sigma"],
                                            counts=True)
        assert abs(k - c["K"]) <= 4e-15 * abs(c["K"]), (c["name"], k, c["K"])
        assert npairs == c["n_internal"][0] * c["n_internal"][1]
        if max(c["n_internal"]) <= 120:
            assert (npairs, m, r) == po.pair_counts(a, b)


def test_kernel_is_symmetric_up_to_rounding(kat):
    c = {x["name"]: x for x in kat["cases"]}
    assert abs(c["n100_vs_n300"]["K"] - c["n300_vs_n100"]["K"]) < 1e-12 * c["n100_vs_n300"]["K"]


def test_labelled_reduces_to_unlabelled_with_one_state():
    from kamphir_b200 import synth
    a = po.prep_kamphir(po.read_newick(synth.kingman_newick(40, 5, states=1)), "mean", po.state_from_name)
    b = po.prep_kamphir(po.read_newick(synth.yule_newick(35, 6, states=1)), "mean", po.state_from_name)
    assert po.kernel_arrays_labelled(a, b, 0.2, 2.0) == po.kernel_arrays(a, b, 0.2, 2.0)


def test_kamphir_compute_restatement():
    """kamphir.py:255-306 restated: identical trees give kcoal = knorm = 1."""
    from kamphir_b200 import synth
    nwk = synth.kingman_newick(30, 3)
    mean, scores = po.kamphir_evaluate([nwk, nwk], nwk)
    assert abs(mean - 1.0) < 1e-12 and all(abs(s - 1.0) < 1e-12 for s in scores)
    other = synth.kingman_newick(30, 4)
    mean2, _ = po.kamphir_evaluate([other], nwk)
    assert 0.0 < mean2 < 1.0


def test_shipped_reference_module_reproduces_the_goldens(kat):
    """oracle/_ref/phyloK2.py (the reference's own file, placed by oracle/Makefile; what bench.py times as the CPU
    baseline) gives the reference's literals and the committed KAT bit for bit.  Skipped where the recipe has not run."""
    from oracle import ref_runner as rr
    if not rr.available():
        pytest.skip("oracle/_ref/phyloK2.py not built (no /root/reference here)")
    from io import StringIO
    from Bio import Phylo
    pk = rr.make_kernel(0.5, 1.0, 1.0)
    t1 = Phylo.read(StringIO("( ( A:0.5, B:0.25 )E:0.5, ( C:0.25, D:0.25 )F:0.5 )G;"), "newick")
    t2 = Phylo.read(StringIO("( ( ( A:0.25, B:0.25 )E:0.5, C:0.25 )F:0.5, D:0.25 )G;"), "newick")
    assert pk.kernel(t1, t2) == 1.125 * (1 + math.exp(-0.0625))                  # tests/unittests.py:95
    done = 0
    for case in kat["cases"]:
        if case["prep"] != "kamphir:mean" or done >= 6:
            continue
        a, b = case_newick(case["a"]), case_newick(case["b"])
        if max(a.count(","), b.count(",")) > 320:
            continue
        pk = rr.make_kernel(case["lam"], case["tau"], case["sigma"])
        k = pk.kernel(rr.prepare(pk, a), rr.prepare(pk, b))
        assert k == case["K"], case["name"]
        done += 1
    assert done >= 3


def test_recursive_definition_agrees_with_the_reference_literal_and_the_dp():
    """oracle.kernel_recursive: the kernel straight from its recursive definition on tree objects (no DP table).  It
    reproduces the reference's closed-form unit-test value and the DP restatement on random small trees."""
    from kamphir_b200 import synth
    k = po.kernel_recursive(po.read_newick(T1), po.read_newick(T2), 0.5, 1.0)
    assert abs(k - 1.125 * (1 + math.exp(-0.0625))) <= 1e-15 * k                 # tests/unittests.py:95
    for seed in range(6):
        a = synth.kingman_newick(12 + 3 * seed, 100 + seed)
        b = synth.yule_newick(30 - 2 * seed, 200 + seed)
        ta, tb = po.read_newick(a), po.read_newick(b)
        fa, fb = po.prep_kamphir(ta), po.prep_kamphir(tb)        # prepares ta, tb in place
        want = po.kernel_arrays(fa, fb, 0.2, 2.0, 1.0)
        got = po.kernel_recursive(ta, tb, 0.2, 2.0, 1.0)
        assert abs(got - want) <= 1e-13 * want


@pytest.mark.parametrize("S,prev", [(2, [0.5, 0.5]), (3, [0.2, 0.5, 0.3])])
def test_labelled_extension_against_the_recursive_definition(S, prev):
    """The tip-state mode has no reference implementation: its DP restatement is pinned as far as it can be,
    against the independent recursive definition, on <= 30-tip trees with the DiffRisk / Stages state counts."""
    from kamphir_b200 import synth
    for seed in range(5):
        a = synth.kingman_newick(18 + 2 * seed, 300 + seed, states=S, prevalence=prev)
        b = synth.yule_newick(30 - seed, 400 + seed, states=S, prevalence=prev)
        ta, tb = po.read_newick(a), po.read_newick(b)
        fa = po.prep_kamphir(ta, "mean", po.state_from_name)
        fb = po.prep_kamphir(tb, "mean", po.state_from_name)
        for x, y, tx, ty in ((fa, fb, ta, tb), (fa, fa, ta, ta)):
            want = po.kernel_arrays_labelled(x, y, 0.2, 2.0, 1.0)
            got = po.kernel_recursive(tx, ty, 0.2, 2.0, 1.0, states=po.state_from_name)
            assert want > 0 and abs(got - want) <= 1e-13 * want
        # the C restatement fed with classes (what the GPU tests and bench.py check the labelled mode against)
        ca, cb = c_oracle.labelled_cflat(fa, S), c_oracle.labelled_cflat(fb, S)
        assert abs(c_oracle.kernel(ca, cb, 0.2, 2.0) - po.kernel_arrays_labelled(fa, fb, 0.2, 2.0)) <= 1e-13 * want
        # matching on is a restriction of matching off: fewer node pairs match, every C is smaller or equal
        assert po.kernel_arrays_labelled(fa, fb, 0.2, 2.0) <= po.kernel_arrays(fa, fb, 0.2, 2.0) * (1 + 1e-12)
