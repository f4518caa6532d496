"""Reading source trees: ``load_trees`` (ref: src/sc_supertree/load.py:7-23) and its flat twin ``load_forest``."""

from __future__ import annotations

import os
from pathlib import Path

from .tree import PhyloNode, make_tree


def load_trees(source_tree_file: str | os.PathLike) -> list[PhyloNode]:
    """The trees of a file that holds one Newick string per line, in file order.

    Same contract as the reference's ``load_trees``: every line is stripped and parsed on its own (an empty
    line is a syntax error), and the result is a list of tree objects ready for ``construct_supertree``.
    """
    trees: list[PhyloNode] = []
    with Path(source_tree_file).open() as handle:
        for line in handle:
            trees.append(make_tree(line.strip()))
    return trees


def load_forest(source_tree_file: str | os.PathLike):
    """The same file as ``load_trees`` reads, parsed natively into the flat source-tree store
    (``engine.Forest``, ``scs_forest_parse_newick``) without building a node object per tree node: what the
    ``scs`` command uses, since at 10 000 taxa x 1 000 trees the node objects cost more than the supertree.
    Feed the result to ``scs.supertree_of_forest``."""
    from .engine import Forest

    return Forest.from_newick(Path(source_tree_file).read_bytes())
