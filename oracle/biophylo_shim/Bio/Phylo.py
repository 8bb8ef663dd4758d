"""Stand-in for the handful of ``Bio.Phylo`` features the reference touches.

TEST INFRASTRUCTURE ONLY.  Biopython is not installed in this image (and there is no network), but the
reference kernel file ``phyloK2.py`` only needs ``from Bio import Phylo`` plus a few ``Tree``/``Clade``
methods.  With this directory on ``sys.path`` the reference file imports *unchanged* under CPython 3.12,
which is how ``tests/golden/make_golden.py`` produces the committed known-answer vectors.

Behaviour mirrored (documented Biopython semantics; Biopython itself is un-vendored and un-pinned in the
reference, ``README.md:15``):

* Newick parsing: children kept in input order, ``branch_length`` is ``None`` when absent, single-quoted
  labels, ``[comments]`` dropped, several trees per handle split on ``;``; the resulting ``Tree`` has
  ``rooted=False`` (call sites: ``kamphir.py:112,323,373``; ``tests/unittests.py:89-90,99``).
* ``Tree.ladderize()``: recursive, ascending by terminal count, *stable* on ties (Python ``list.sort``).
* ``get_terminals/get_nonterminals/find_clades(order=...)``: depth-first pre-order (default) / post-order.
* ``total_branch_length()``: sum over clades whose branch length is truthy, in pre-order.  Summed
  left-to-right with plain ``+`` (CPython 2.7 behaviour, the interpreter the reference targets); CPython
  3.12's built-in ``sum`` is compensated and could differ in the last ulp.
* ``depths()``: root depth is ``root.branch_length or 0``.

Nothing under ``kamphir_b200/`` imports this module.
"""

__all__ = ["read", "parse", "Tree", "Clade", "NewickError"]


class NewickError(Exception):
    pass


class Clade(object):
    def __init__(self, branch_length=None, name=None, clades=None):
        self.branch_length = branch_length
        self.name = name
        self.clades = clades if clades is not None else []
        self.confidence = None
        self.comment = None

    # -- structure -----------------------------------------------------------------------------------
    @property
    def root(self):
        return self

    def is_terminal(self):
        return not self.clades

    def _preorder(self):
        stack = [self]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.clades))

    def _postorder(self):
        # iterative post-order: children (in order) before their parent
        out = []
        stack = [(self, 0)]
        while stack:
            node, i = stack.pop()
            if i < len(node.clades):
                stack.append((node, i + 1))
                stack.append((node.clades[i], 0))
            else:
                out.append(node)
        return out

    def find_clades(self, terminal=None, order="preorder", **kwargs):
        if kwargs:
            raise NotImplementedError("shim: attribute filters %r" % (kwargs,))
        if order == "preorder":
            it = self._preorder()
        elif order == "postorder":
            it = iter(self._postorder())
        else:
            raise ValueError("shim supports preorder/postorder only")
        for node in it:
            if terminal is None or node.is_terminal() == terminal:
                yield node

    def get_terminals(self, order="preorder"):
        return list(self.find_clades(terminal=True, order=order))

    def get_nonterminals(self, order="preorder"):
        return list(self.find_clades(terminal=False, order=order))

    def count_terminals(self):
        return sum(1 for _ in self.find_clades(terminal=True))

    def total_branch_length(self):
        total = 0
        for node in self._preorder():
            if node.branch_length:
                total = total + node.branch_length
        return total

    def depths(self, unit_branch_lengths=False):
        depths = {}
        stack = [(self, (1 if unit_branch_lengths else (self.branch_length or 0)))]
        while stack:
            node, d = stack.pop()
            depths[node] = d
            for child in node.clades:
                stack.append((child, d + (1 if unit_branch_lengths else (child.branch_length or 0))))
        return depths

    def ladderize(self, reverse=False):
        # iterative form of Biopython's recursive definition (deep caterpillars exceed the recursion limit)
        counts = {}
        for node in self._postorder():
            counts[id(node)] = 1 if not node.clades else sum(counts[id(c)] for c in node.clades)
        for node in self._preorder():
            node.clades.sort(key=lambda c: counts[id(c)], reverse=reverse)

    def __str__(self):
        return self.name or self.__class__.__name__


class Tree(object):
    def __init__(self, root=None, rooted=False, name=None):
        self.root = root if root is not None else Clade()
        self.rooted = rooted
        self.name = name

    @property
    def clade(self):
        return self.root

    def find_clades(self, *a, **kw):
        return self.root.find_clades(*a, **kw)

    def get_terminals(self, order="preorder"):
        return self.root.get_terminals(order)

    def get_nonterminals(self, order="preorder"):
        return self.root.get_nonterminals(order)

    def count_terminals(self):
        return self.root.count_terminals()

    def total_branch_length(self):
        return self.root.total_branch_length()

    def depths(self, unit_branch_lengths=False):
        return self.root.depths(unit_branch_lengths)

    def ladderize(self, reverse=False):
        self.root.ladderize(reverse)

    def format(self, fmt="newick"):
        return to_newick(self)


# ---------------------------------------------------------------------------------------------------
# Newick reader / writer
# ---------------------------------------------------------------------------------------------------
_PUNCT = "()[]':;,"


def _split_trees(text):
    """Split a handle's text into one string per tree on ';' outside quotes/comments."""
    out, buf, i, n = [], [], 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "'":
            j = text.find("'", i + 1)
            if j < 0:
                raise NewickError("unterminated quoted label")
            buf.append(text[i:j + 1])
            i = j + 1
            continue
        if ch == "[":
            j = text.find("]", i + 1)
            if j < 0:
                raise NewickError("unterminated comment")
            i = j + 1
            continue
        if ch == ";":
            s = "".join(buf).strip()
            if s:
                out.append(s)
            buf = []
        else:
            buf.append(ch)
        i += 1
    s = "".join(buf).strip()
    if s:
        out.append(s)
    return out


def _parse_one(text):
    root = Clade()
    current = root
    parents = {}
    i, n = 0, len(text)
    lp = rp = 0
    synthetic_root = False
    while i < n:
        ch = text[i]
        if ch.isspace():
            i += 1
        elif ch == "(":
            child = Clade()
            parents[id(child)] = current
            current.clades.append(child)
            current = child
            lp += 1
            i += 1
        elif ch == ",":
            parent = parents.get(id(current))
            if parent is None:
                # missing outer parentheses: make a new root above the current clade
                new_root = Clade()
                new_root.clades.append(current)
                parents[id(current)] = new_root
                root = new_root
                parent = new_root
                synthetic_root = True
            child = Clade()
            parents[id(child)] = parent
            parent.clades.append(child)
            current = child
            i += 1
        elif ch == ")":
            parent = parents.get(id(current))
            if parent is None:
                raise NewickError("Parenthesis mismatch.")
            current = parent
            rp += 1
            i += 1
        elif ch == ":":
            j = i + 1
            while j < n and text[j] == " ":
                j += 1
            k = j
            while k < n and text[k] not in _PUNCT and not text[k].isspace():
                k += 1
            try:
                current.branch_length = float(text[j:k])
            except ValueError:
                raise NewickError("bad branch length %r" % text[j:k])
            i = k
        elif ch == "'":
            j = text.find("'", i + 1)
            current.name = text[i + 1:j]
            i = j + 1
        else:
            k = i
            while k < n and text[k] not in _PUNCT and not text[k].isspace():
                k += 1
            if k == i:
                raise NewickError("unexpected character %r" % ch)
            current.name = text[i:k]
            i = k
    if lp != rp:
        raise NewickError("Number of open/close parentheses do not match.")
    if current is not root and not (synthetic_root and parents.get(id(current)) is root):
        raise NewickError("Parenthesis mismatch.")
    return Tree(root=root, rooted=False)


def _text_of(handle):
    if hasattr(handle, "read"):
        return handle.read()
    with open(handle) as fh:
        return fh.read()


def parse(handle, fmt="newick", **kwargs):
    if fmt != "newick":
        raise ValueError("shim supports newick only")
    for chunk in _split_trees(_text_of(handle)):
        yield _parse_one(chunk)


def read(handle, fmt="newick", **kwargs):
    trees = list(parse(handle, fmt))
    if len(trees) != 1:
        raise ValueError("expected exactly one tree, found %d" % len(trees))
    return trees[0]


def to_newick(tree, fmt="%.17g"):
    def rec(node):
        s = ""
        if node.clades:
            s = "(" + ",".join(rec(c) for c in node.clades) + ")"
        if node.name:
            s += node.name
        if node.branch_length is not None:
            s += ":" + (fmt % node.branch_length)
        return s
    return rec(tree.root) + ";"
