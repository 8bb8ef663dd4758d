"""Seeded synthetic phylogenies for tests and benchmarks (SURVEY.md 8d "Synthetic inputs").

* Kingman coalescent -- the process behind ``ape::rcoal`` that made the reference's ``tests/test.tree``
  (``tests/make_test_data.R:5-9``): n lineages; while k > 1 wait Exp(k(k-1)/2) and merge a uniform random pair.
* Yule -- rate-1 pure birth from one lineage until n tips, then one more Exp(n) increment on the pendant edges.

Trees are returned as Newick strings with 10 significant digits per branch length (what rcolgem's
``write.tree`` emits, ``tests/unittests.py:141``), children in generation order, so every consumer (the C++
flattener, the oracle, the reference itself) sees byte-identical input.  Multi-type tips are named
``"<i>_<state>"`` like ``rcolgem.py:274``.
"""
import numpy as np

__all__ = ["kingman_newick", "yule_newick", "caterpillar_newick", "balanced_newick", "tree_set"]


def _emit(children, blen, names, root):
    """Iterative Newick writer over arrays (children[v] = (l, r) or None)."""
    out = []
    stack = [(root, 0)]
    while stack:
        v, st = stack.pop()
        kids = children[v]
        if kids is None:
            out.append(names[v])
            if v != root:
                out.append(":%.10g" % blen[v])
            continue
        if st == 0:
            out.append("(")
            stack.append((v, 1))
            stack.append((kids[0], 0))
        elif st == 1:
            out.append(",")
            stack.append((v, 2))
            stack.append((kids[1], 0))
        else:
            out.append(")")
            if v != root:
                out.append(":%.10g" % blen[v])
    out.append(";")
    return "".join(out)


def _tip_names(n, rng, states, prevalence):
    if not states:
        return ["t%d" % (i + 1) for i in range(n)]
    p = prevalence if prevalence is not None else [1.0 / states] * states
    st = rng.choice(states, size=n, p=p)
    return ["%d_%d" % (i + 1, st[i]) for i in range(n)]


def kingman_newick(n, seed, states=0, prevalence=None):
    """Kingman coalescent tree with n tips (ultrametric)."""
    rng = np.random.default_rng(seed)
    total = 2 * n - 1
    children = [None] * total
    height = [0.0] * total
    blen = [0.0] * total
    names = _tip_names(n, rng, states, prevalence) + [None] * (n - 1)
    waits = rng.exponential(size=n - 1)
    picks = rng.random((n - 1, 2))
    active = list(range(n))
    t = 0.0
    nxt = n
    for step in range(n - 1):
        k = len(active)
        t += waits[step] * (2.0 / (k * (k - 1)))
        i = int(picks[step, 0] * k)
        j = int(picks[step, 1] * (k - 1))
        if j >= i:
            j += 1
        a, b = active[i], active[j]
        children[nxt] = (a, b)
        height[nxt] = t
        blen[a] = t - height[a]
        blen[b] = t - height[b]
        hi, lo = (i, j) if i > j else (j, i)
        active[hi] = active[-1]
        active.pop()
        if lo < len(active):
            active[lo] = nxt
        else:                       # lo was the last slot and has just been popped (cannot happen: lo < hi)
            active.append(nxt)
        nxt += 1
    return _emit(children, blen, names, active[0])


def yule_newick(n, seed, states=0, prevalence=None):
    """Yule (pure-birth, rate 1) tree with n tips."""
    rng = np.random.default_rng(seed)
    total = 2 * n - 1
    children = [None] * total
    birth = [0.0] * total
    blen = [0.0] * total
    root = 0
    leaves = [0]
    waits = rng.exponential(size=n)
    picks = rng.random(n)
    t = 0.0
    nxt = 1
    step = 0
    while len(leaves) < n:
        k = len(leaves)
        t += waits[step] / k
        idx = int(picks[step] * k)
        step += 1
        v = leaves[idx]
        blen[v] = t - birth[v]
        children[v] = (nxt, nxt + 1)
        birth[nxt] = birth[nxt + 1] = t
        leaves[idx] = nxt
        leaves.append(nxt + 1)
        nxt += 2
    t += waits[n - 1] / n
    for v in leaves:
        blen[v] = t - birth[v]
    tipnames = _tip_names(n, rng, states, prevalence)
    names = [None] * total
    for i, v in enumerate(sorted(leaves)):
        names[v] = tipnames[i]
    return _emit(children, blen, names, root)


def caterpillar_newick(n, seed=0):
    """Fully unbalanced tree (worst case for the level wavefront: n-1 levels)."""
    rng = np.random.default_rng(seed)
    s = "(t1:%.10g,t2:%.10g)" % (rng.exponential(), rng.exponential())
    parts = [s]
    for i in range(3, n + 1):
        parts.insert(0, "(")
        parts.append(":%.10g,t%d:%.10g)" % (rng.exponential(), i, rng.exponential()))
    return "".join(parts) + ";"


def balanced_newick(depth, seed=0):
    """Perfectly balanced tree with 2**depth tips."""
    rng = np.random.default_rng(seed)
    level = ["t%d" % (i + 1) for i in range(2 ** depth)]
    while len(level) > 1:
        level = ["(%s:%.10g,%s:%.10g)" % (level[i], rng.exponential(), level[i + 1], rng.exponential())
                 for i in range(0, len(level), 2)]
    return level[0] + ";"


def tree_set(n_trees, n_tips, base_seed, kind="mixed", states=0, prevalence=None):
    """n_trees Newick strings; kind in {'kingman','yule','mixed'} (mixed = first half Kingman, rest Yule)."""
    out = []
    for i in range(n_trees):
        seed = base_seed + i
        if kind == "kingman" or (kind == "mixed" and i < (n_trees + 1) // 2):
            out.append(kingman_newick(n_tips, seed, states, prevalence))
        else:
            out.append(yule_newick(n_tips, seed, states, prevalence))
    return out


# ---------------------------------------------------------------------------------------------------
# a minimal Bio.Phylo-shaped tree (what PhyloKernel's object path needs), for benchmarks and examples that must not
# depend on Biopython: .root, .rooted, .clades, .branch_length, .name, get_terminals(), get_nonterminals(order)
# ---------------------------------------------------------------------------------------------------
class Clade(object):
    def __init__(self, branch_length=None, name=None):
        self.branch_length = branch_length
        self.name = name
        self.clades = []


class Tree(object):
    rooted = False

    def __init__(self, root):
        self.root = root

    def _walk(self, post):
        out, stack = [], [(self.root, False)]
        while stack:
            node, done = stack.pop()
            if done or not node.clades:
                out.append(node)
                continue
            if post:
                stack.append((node, True))
                stack.extend((c, False) for c in reversed(node.clades))
            else:
                out.append(node)
                stack.extend((c, False) for c in reversed(node.clades))
        return out

    def get_terminals(self, order="preorder"):
        return [n for n in self._walk(order == "postorder") if not n.clades]

    def get_nonterminals(self, order="preorder"):
        return [n for n in self._walk(order == "postorder") if n.clades]

    def total_branch_length(self):
        return sum(n.branch_length for n in self._walk(False) if n.branch_length)


def newick_to_tree(text):
    """Parse one Newick string of the dialect this module writes into a Tree of Clade objects."""
    root = Clade()
    node, stack, i, n = root, [], 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "(":
            child = Clade()
            node.clades.append(child)
            stack.append(node)
            node = child
            i += 1
        elif ch == ",":
            node = Clade()
            stack[-1].clades.append(node)
            i += 1
        elif ch == ")":
            node = stack.pop()
            i += 1
        elif ch == ";":
            break
        elif ch == ":":
            j = i + 1
            while j < n and text[j] not in ",();":
                j += 1
            node.branch_length = float(text[i + 1:j])
            i = j
        elif ch.isspace():
            i += 1
        else:
            j = i
            while j < n and text[j] not in ":,();":
                j += 1
            node.name = text[i:j]
            i = j
    return Tree(root)
