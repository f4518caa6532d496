"""Reading source trees: ``load_trees`` (ref: src/sc_supertree/load.py:7-23) and its flat twin ``load_forest``."""

from __future__ import annotations

import os
from collections.abc import Sequence
from pathlib import Path

from .tree import NewickError, PhyloNode, make_tree

_NODE_SLOTS = frozenset(PhyloNode.__slots__)


class _SourceFile:
    """What ``load_trees`` read: the stripped lines of the file and the forest the native parser made of them
    (``scs_forest_parse_newick``), shared by the tree objects of that call."""

    __slots__ = ("count", "forest", "lines")

    def __init__(self, lines: list[str], forest) -> None:
        self.lines = lines
        self.forest = forest
        self.count = len(lines)


class LoadedTree(PhyloNode):
    """A tree of a loaded file whose node objects are built the first time anything looks at them.

    ``load_trees`` parses the file natively into the flat source-tree store; at 10 000 taxa x 1 000 trees the 1.3
    million node objects of the reference's ``list[PhyloNode]`` cost 10 s of pure-Python parsing that
    ``construct_supertree`` never needs (it flattens them again).  Each element of the returned list is the root of
    its tree: an ordinary ``PhyloNode`` as soon as an attribute or method is used (the line is then parsed by
    ``make_tree``, exactly as before), and only a line number until then.  ``construct_supertree`` recognises a list
    nobody has touched and takes the forest that was parsed natively."""

    __slots__ = ("_index", "_source")

    def __init__(self, source: _SourceFile, index: int) -> None:  # the node slots stay unset until _build
        object.__setattr__(self, "_source", source)
        object.__setattr__(self, "_index", index)

    def _pending(self) -> _SourceFile | None:
        try:
            return object.__getattribute__(self, "_source")
        except AttributeError:  # a copy being rebuilt by pickle / copy: no source yet
            return None

    def _build(self, source: _SourceFile) -> None:
        object.__setattr__(self, "_source", None)
        root = make_tree(source.lines[object.__getattribute__(self, "_index")])
        for slot in ("name", "length", "support"):
            object.__setattr__(self, slot, getattr(root, slot))
        object.__setattr__(self, "parent", None)
        for child in root.children:
            child.parent = self
        object.__setattr__(self, "children", root.children)

    def __getattr__(self, attr: str):
        # reached only when the ordinary lookup failed: an unset node slot of a tree that has not been built
        if attr in _NODE_SLOTS:
            source = self._pending()
            if source is not None:
                self._build(source)
                return object.__getattribute__(self, attr)
        raise AttributeError(attr)

    def __setattr__(self, attr: str, value) -> None:
        if attr in _NODE_SLOTS:
            source = self._pending()
            if source is not None:
                self._build(source)
        object.__setattr__(self, attr, value)


def untouched_forest(trees: Sequence, weights: Sequence[float] | None):
    """The natively parsed forest behind ``trees`` if they are exactly the list one ``load_trees`` call returned and
    nobody has looked inside any of them (so nobody can have changed them); ``None`` otherwise.  With ``weights`` the
    forest is re-made around the same arrays with those tree weights."""
    if len(trees) == 0 or type(trees[0]) is not LoadedTree:
        return None
    source = trees[0]._pending()
    if source is None or len(trees) != source.count:
        return None
    for i, tree in enumerate(trees):
        if type(tree) is not LoadedTree or tree._pending() is not source or object.__getattribute__(tree, "_index") != i:
            return None
    forest = source.forest
    if weights is None or all(float(w) == 1.0 for w in weights):
        return forest
    import numpy as np

    from .engine import Forest

    arrays = [forest.tree_arrays(t) for t in range(forest.num_trees)]
    offsets = np.zeros(len(arrays) + 1, dtype=np.int64)
    np.cumsum([len(a[0]) for a in arrays], out=offsets[1:])
    parent, length, support, taxon = (np.concatenate([a[k] for a in arrays]) for k in range(4))
    return Forest.from_arrays(offsets, parent, length, support, taxon, [float(w) for w in weights], forest.names, copy=True)


def load_trees(source_tree_file: str | os.PathLike) -> list[PhyloNode]:
    """The trees of a file that holds one Newick string per line, in file order.

    Same contract as the reference's ``load_trees``: every line is stripped and parsed on its own (an empty
    line is a syntax error), and the result is a list of tree objects ready for ``construct_supertree``.  The
    file is parsed natively, which also checks its syntax; the node objects of a tree are built when something
    first looks at that tree (``LoadedTree``).
    """
    with Path(source_tree_file).open() as handle:
        lines = [line.strip() for line in handle]
    if not lines:
        return []
    from .engine import Forest

    try:
        forest = Forest.from_newick("\n".join(lines) + "\n")
    except NewickError:
        # a syntax error (make_tree says which, in its own words) or trees the flat store refuses -- a taxon on
        # two tips of one tree -- but the reference's loader accepts: node objects right away, as the reference does
        return [make_tree(line) for line in lines]
    if forest.num_trees != len(lines):  # cannot happen: one tree per line
        return [make_tree(line) for line in lines]
    source = _SourceFile(lines, forest)
    return [LoadedTree(source, i) for i in range(len(lines))]


def load_forest(source_tree_file: str | os.PathLike):
    """The same file as ``load_trees`` reads, parsed natively into the flat source-tree store
    (``engine.Forest``, ``scs_forest_parse_newick``) without building a node object per tree node: what the
    ``scs`` command uses, since at 10 000 taxa x 1 000 trees the node objects cost more than the supertree.
    Feed the result to ``scs.supertree_of_forest``."""
    from .engine import Forest

    return Forest.from_newick(Path(source_tree_file).read_bytes())
