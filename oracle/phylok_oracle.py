"""CPU oracle for the phyloK2 subset-tree kernel path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  Nothing under ``kamphir_b200/`` does; the product path fails loudly without its CUDA
library.

It restates, in plain Python over flat per-node arrays, the arithmetic of the reference
(``/root/reference``; citations are ``file:line`` relative to that tree):

=====================  ==========================================================================
``normalize_tree``     ``phyloK2.py:101-130``  (mean / median rescale of every non-root branch)
``annotate``           ``phyloK2.py:132-146``  (production = #tip children + 1, post-order index, bl, sqbl)
``kernel_arrays``      ``phyloK2.py:158-209``  (the O(n1*n2) node-pair DP; the hot path)
``prep_kamphir``       ``kamphir.py:153-156`` and ``kamphir.py:279-282``  (root bl 0 -> ladderize -> normalise -> annotate)
``kamphir_compute``    ``kamphir.py:255-306``  (kcoal cosine of sorted node heights, log-space normalisation)
``kamphir_evaluate``   ``kamphir.py:434-453``  (mean over replicates)
``gram``               ``phyloK2.py:148-156``  (intent: symmetric matrix of raw K, j >= i)
=====================  ==========================================================================

Parity pin: ``tests/test_oracle.py`` checks this file against (a) the reference's own golden values
(``tests/unittests.py:93-113``) and (b) ``tests/golden/kernel_kat.json``, produced in the build container by
``tests/golden/make_golden.py``, which imports the reference's *unmodified* ``phyloK2.py`` through the
``Bio.Phylo`` stand-in in ``oracle/biophylo_shim``.  The tip-state ("labelled") mode at the bottom has no
reference implementation and no reference test: parity unpinned for that mode (it must reduce to the
unlabelled kernel when all states are equal, which *is* tested).

The tree container is the stand-in ``Bio.Phylo`` of ``oracle/biophylo_shim`` (Biopython is not installed here).
"""
import math
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "biophylo_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from Bio import Phylo  # noqa: E402  (the stand-in)
from io import StringIO  # noqa: E402


# ---------------------------------------------------------------------------------------------------
# tree helpers
# ---------------------------------------------------------------------------------------------------
def read_newick(text):
    """One tree from a Newick string (``kamphir.py:323``)."""
    return Phylo.read(StringIO(text), "newick")


def parse_newick_file(path):
    """All trees of a Newick file (``kamphir.py:112``, ``tests/unittests.py:99``)."""
    return list(Phylo.parse(path, "newick"))


def normalize_tree(tree, mode="median"):
    """``phyloK2.py:101-130``.  Root is skipped because Newick trees are ``rooted=False`` (``:116``)."""
    branches = tree.get_nonterminals() + tree.get_terminals()          # pre-order, internal first  (:108)
    nbranches = len(branches) - 1                                      # (:109)
    non_root = branches[int(not tree.rooted):]
    if mode == "mean":
        tree_length = tree.total_branch_length() - tree.root.branch_length   # (:113)
        scale = tree_length / nbranches
    elif mode == "median":
        lengths = sorted(b.branch_length for b in non_root)             # (:120-121)
        half = nbranches // 2                                           # py2 integer '/' (:123-127)
        if nbranches % 2 == 0:
            scale = (lengths[half - 1] + lengths[half]) / 2.
        else:
            scale = lengths[half]
    else:
        return                                                          # 'none' matches neither branch
    for b in non_root:
        b.branch_length /= scale


class Flat(object):
    """Post-order per-internal-node arrays: what ``annotate_tree`` hangs on the Clade objects."""
    __slots__ = ("n", "prod", "cprod", "cidx", "bl", "sqbl", "ntips", "state")

    def __init__(self, n):
        self.n = n
        self.prod = [0] * n           # node.production  (1..3)
        self.cprod = [None] * n       # production of each child (0 = tip)
        self.cidx = [None] * n        # post-order index of each internal child, -1 for a tip
        self.bl = [None] * n          # node.bl   = child branch lengths
        self.sqbl = [0.0] * n         # node.sqbl = sum of squares
        self.ntips = 0
        self.state = None             # labelled mode only: per node, per child, tip-state id or -1


def annotate(tree, states=None):
    """``phyloK2.py:132-146`` as arrays.  ``states``: optional callable tip-name -> small int (extension)."""
    nodes = tree.get_nonterminals(order="postorder")
    index = {}
    for i, node in enumerate(nodes):
        index[id(node)] = i
    flat = Flat(len(nodes))
    flat.ntips = len(tree.get_terminals())
    if states is not None:
        flat.state = [None] * flat.n
    for i, node in enumerate(nodes):
        kids = node.clades
        if len(kids) != 2:
            # reference: IndexError at phyloK2.py:188-189 (unary) / silently ignores 3rd+ child (polytomy)
            raise ValueError("node %d has %d children; the kernel is defined for strictly binary trees"
                             % (i, len(kids)))
        kp = [0 if not c.clades else None for c in kids]
        flat.cidx[i] = [-1 if not c.clades else index[id(c)] for c in kids]
        for s in (0, 1):
            if kp[s] is None:
                kp[s] = flat.prod[flat.cidx[i][s]]
        flat.cprod[i] = kp
        flat.prod[i] = sum(1 for p in kp if p == 0) + 1                 # (:141-142)
        lens = [c.branch_length for c in kids]                          # (:144)
        flat.bl[i] = lens
        flat.sqbl[i] = sum([b ** 2 for b in lens])                      # (:146)
        if states is not None:
            flat.state[i] = [states(c.name) if not c.clades else -1 for c in kids]
    return flat


def prep_kamphir(tree, mode="mean", states=None):
    """``kamphir.py:153-156`` / ``:279-282``: root bl 0 -> ladderize -> normalise -> annotate."""
    tree.root.branch_length = 0.
    tree.ladderize()
    normalize_tree(tree, mode)
    return annotate(tree, states)


# ---------------------------------------------------------------------------------------------------
# the hot path
# ---------------------------------------------------------------------------------------------------
def kernel_arrays(A, B, decayFactor=0.1, gaussFactor=1., sigma=1.):
    """``phyloK2.py:158-209``: K = sum over production-matched internal-node pairs of C(n1,n2).

    Same loop nest (post-order x post-order), same expanded Gaussian argument ``sq1 + sq2 - 2*dot``
    (``:184-185``), same ordered-slot child rule (``:187-199``).  ``dp`` is a dense list-of-lists instead of
    the reference's dict; the KeyError branch (``:200-201``) is unreachable outside the broken striped mode.
    """
    lam = decayFactor
    k = 0
    dp = [None] * A.n
    b_prod, b_cprod, b_cidx, b_bl, b_sqbl = B.prod, B.cprod, B.cidx, B.bl, B.sqbl
    rng_b = range(B.n)
    for i in range(A.n):
        p1 = A.prod[i]
        cp1 = A.cprod[i]
        ci1 = A.cidx[i]
        bl1 = A.bl[i]
        sq1 = A.sqbl[i]
        row = [0.0] * B.n
        dp[i] = row
        for j in rng_b:
            if p1 != b_prod[j]:
                continue
            bl2 = b_bl[j]
            res = lam * math.exp(-1. / gaussFactor * (sq1 + b_sqbl[j] - 2 * (bl1[0] * bl2[0] + bl1[1] * bl2[1])))
            cp2 = b_cprod[j]
            for s in (0, 1):
                if cp1[s] != cp2[s]:
                    continue
                if cp1[s] == 0:
                    res *= sigma + lam
                else:
                    res *= sigma + dp[ci1[s]][b_cidx[j][s]]
            row[j] = res
            k += res
    return k


def kernel_arrays_labelled(A, B, decayFactor=0.1, gaussFactor=1., sigma=1.):
    """Tip-state extension (no reference implementation; SURVEY 8c 'parity unpinned').

    Two nodes match when they have the same production AND the same ordered tuple of tip-child states; a
    tip/tip child slot contributes (sigma + lambda) only when the two tips carry the same state.  With a
    single state everywhere this is exactly ``kernel_arrays``.
    """
    lam = decayFactor
    k = 0
    dp = [None] * A.n
    for i in range(A.n):
        row = [None] * B.n
        dp[i] = row
        for j in range(B.n):
            if A.prod[i] != B.prod[j] or A.state[i] != B.state[j]:
                continue
            bl1, bl2 = A.bl[i], B.bl[j]
            res = lam * math.exp(-1. / gaussFactor * (A.sqbl[i] + B.sqbl[j]
                                                      - 2 * (bl1[0] * bl2[0] + bl1[1] * bl2[1])))
            for s in (0, 1):
                if A.cprod[i][s] != B.cprod[j][s]:
                    continue
                if A.cprod[i][s] == 0:
                    res *= sigma + lam
                else:
                    c = dp[A.cidx[i][s]][B.cidx[j][s]]
                    if c is not None:                      # children matched on states too
                        res *= sigma + c
            row[j] = res
            k += res
    return k


def kernel_recursive(t1, t2, decayFactor=0.1, gaussFactor=1., sigma=1., states=None):
    """The subset-tree kernel written down a second, independent way: straight from its recursive definition on the tree
    OBJECTS, with no DP table, no flat arrays and no post-order bookkeeping (exponential in nothing, but O(n1 n2 depth):
    for trees of a few dozen tips).  It is the checker of ``kernel_arrays`` and, above all, of the tip-state extension
    ``kernel_arrays_labelled``, which has no reference implementation to be compared with.

      K(t1, t2) = sum over internal n1 of t1, internal n2 of t2 of C(n1, n2)
      C(n1, n2) = undefined (no match)            if the productions of n1 and n2 differ
                = lambda * exp(-|bl(n1) - bl(n2)|^2 / tau) * prod over the child slots s = 0, 1 of
                      sigma + lambda              both children are tips
                      sigma + C(c1_s, c2_s)       both are internal and C of them is defined
                      1                           otherwise
      production of a node = number of its tip children (phyloK2.py:141-142), and, with ``states`` (tip name -> state),
      the ordered tuple of the states of its tip children.
    The Gaussian is written in the difference form (not the reference's expanded one), which is one more difference
    between this function and the ones it checks; agreement is to rounding, not to the bit.
    """
    lam = decayFactor

    def production(node):
        tips = tuple(not c.clades for c in node.clades)
        if states is None:
            return sum(tips)
        return (sum(tips), tuple(states(c.name) if not c.clades else None for c in node.clades))

    def c_value(n1, n2):
        if len(n1.clades) != 2 or len(n2.clades) != 2:
            raise ValueError("strictly binary trees only")
        if production(n1) != production(n2):
            return None
        d0 = n1.clades[0].branch_length - n2.clades[0].branch_length
        d1 = n1.clades[1].branch_length - n2.clades[1].branch_length
        res = lam * math.exp(-(d0 * d0 + d1 * d1) / gaussFactor)
        for s in (0, 1):
            c1, c2 = n1.clades[s], n2.clades[s]
            if not c1.clades and not c2.clades:
                res *= sigma + lam
            elif c1.clades and c2.clades:
                sub = c_value(c1, c2)
                if sub is not None:
                    res *= sigma + sub
        return res

    k = 0.0
    for n1 in t1.get_nonterminals():
        for n2 in t2.get_nonterminals():
            c = c_value(n1, n2)
            if c is not None:
                k += c
    return k


def kernel(t1, t2, decayFactor=0.1, gaussFactor=1., sigma=1.):
    """Tree-object front end: annotate on the fly like ``phyloK2.py:169-172``."""
    return kernel_arrays(annotate(t1), annotate(t2), decayFactor, gaussFactor, sigma)


def pair_counts(A, B):
    """Exact per-DP work counts used by the roofline: node pairs, matched cells M, child reads R."""
    m = r = 0
    for i in range(A.n):
        for j in range(B.n):
            if A.prod[i] != B.prod[j]:
                continue
            m += 1
            for s in (0, 1):
                if A.cprod[i][s] == B.cprod[j][s] and A.cprod[i][s] != 0:
                    r += 1
    return A.n * B.n, m, r


# ---------------------------------------------------------------------------------------------------
# the caller's arithmetic (kamphir.py)
# ---------------------------------------------------------------------------------------------------
def node_heights(tree):
    """``kamphir.py:255-258`` (and ``:118-149`` with tscale=1): ascending heights of internal nodes above the
    deepest tip, from the tree *before* it is normalised."""
    depths = tree.depths()
    height = max(depths.values())
    hs = [height - depths[node] for node in tree.get_nonterminals()]
    hs.sort()
    return height, hs


def kcoal(sim_heights, target_heights):
    """``kamphir.py:260-269``: cosine similarity of the two ascending height vectors, indexed by the
    simulated tree's node rank (IndexError if it has more internal nodes than the target)."""
    dot = n1 = n2 = 0
    for rank, h in enumerate(sim_heights):
        t = target_heights[rank]
        dot += h * t
        n1 += h * h
        n2 += t * t
    return dot / (math.sqrt(n1) * math.sqrt(n2))


def knorm(k, ref_denom, tree_denom):
    """``kamphir.py:300``: K / sqrt(K11*K22) evaluated in log space."""
    return math.exp(math.log(k) - 0.5 * (math.log(ref_denom) + math.log(tree_denom)))


def kamphir_compute(sim_tree, target_flat, target_heights, ref_denom, decayFactor=0.2, gaussFactor=2.0,
                    sigma=1., normalize="mean", kadj=1.0):
    """``kamphir.py:249-306`` for one simulated tree (a stand-in ``Tree``; it is modified in place)."""
    _, hs = node_heights(sim_tree)
    kc = kcoal(hs, target_heights)
    flat = prep_kamphir(sim_tree, normalize)
    k = kernel_arrays(target_flat, flat, decayFactor, gaussFactor, sigma)          # (:290)
    d = kernel_arrays(flat, flat, decayFactor, gaussFactor, sigma)                 # (:291)
    return knorm(k, ref_denom, d) * math.pow(kc, kadj)                             # (:300-306)


def kamphir_evaluate(sim_newicks, target_newick, **kw):
    """``kamphir.py:434-453`` for a single target tree: mean of the per-replicate scores."""
    target = read_newick(target_newick)
    _, t_heights = node_heights(target)                                            # (:118,147-149)
    t_flat = prep_kamphir(target, kw.get("normalize", "mean"))
    ref_denom = kernel_arrays(t_flat, t_flat, kw.get("decayFactor", 0.2), kw.get("gaussFactor", 2.0),
                              kw.get("sigma", 1.))                                 # (:157)
    scores = [kamphir_compute(read_newick(s), t_flat, t_heights, ref_denom, **kw) for s in sim_newicks]
    return sum(scores) / len(scores), scores


def gram(flats, decayFactor=0.1, gaussFactor=1., sigma=1.):
    """``phyloK2.py:148-156`` (intent): symmetric matrix of raw K over j >= i."""
    n = len(flats)
    kmat = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for j in range(i, n):
            kmat[i][j] = kmat[j][i] = kernel_arrays(flats[i], flats[j], decayFactor, gaussFactor, sigma)
    return kmat


def state_from_name(name):
    """Multi-type simulators label tips ``"<i>_<deme>"`` (``rcolgem.py:274,376,588``): state = token after the
    last underscore, mapped through int() when numeric, else hashed by first appearance by the caller."""
    if name is None:
        return 0
    tok = name.rsplit("_", 1)[-1]
    try:
        return int(tok)
    except ValueError:
        return sum(ord(c) for c in tok) & 0xff


def _split_trees_text(text):
    """Newick text of a multi-tree file -> list of single-tree strings (each ending in ';')."""
    return [s + ";" for s in Phylo._split_trees(text)]


def prepare(tree, prep):
    """KAT preparation modes (tests/golden/make_golden.py): 'raw', 'kamphir:<mode>', 'ladder:none' -> Flat."""
    if prep == "raw":
        return annotate(tree)
    kind, mode = prep.split(":")
    if kind == "kamphir":
        return prep_kamphir(tree, mode)
    if kind == "ladder":
        tree.ladderize()
        return annotate(tree)
    raise ValueError(prep)
