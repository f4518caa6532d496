"""``load_trees``: a line-separated Newick file -> list of trees (ref: src/sc_supertree/load.py:7-23)."""

from __future__ import annotations

import os
from pathlib import Path

from .tree import PhyloNode, make_tree


def load_trees(source_tree_file: str | os.PathLike) -> list[PhyloNode]:
    """Load a line-separated file of Newick-formatted trees.

    Parameters
    ----------
    source_tree_file : str | bytes | os.PathLike
        The path to the source tree file.

    Returns
    -------
    list[PhyloNode]
        A list of all source trees in the file.

    """
    with Path(source_tree_file).open() as f:
        return [make_tree(line.strip()) for line in f]


def load_forest(source_tree_file: str | os.PathLike):
    """The same file as ``load_trees`` reads, parsed natively into the flat source-tree store
    (``engine.Forest``, ``scs_forest_parse_newick``) without building a node object per tree node: what the
    ``scs`` command uses, since at 10 000 taxa x 1 000 trees the node objects cost more than the supertree.
    Feed the result to ``scs.supertree_of_forest``."""
    from .engine import Forest

    return Forest.from_newick(Path(source_tree_file).read_bytes())
