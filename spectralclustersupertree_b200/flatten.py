"""Flatten source trees into the leaf-tour arrays the CUDA kernels consume.

One tree becomes its leaves in depth-first order plus, for every pair of consecutive leaves,
the depth and the weighting value of their lowest common ancestor (LCA).  That is all the
proper-cluster-graph build needs: the LCA of any two leaves ``i < j`` of the tour is the
shallowest of the consecutive-leaf LCAs ``i..j-1``, two leaves are a proper cluster iff that
node is not the root (ref: src/sc_supertree/scs.py:570-579,644-658), and the edge weight they
receive is that node's value times the tree weight (ref: scs.py:655-657).

Node values follow the reference's ``length_function`` exactly, evaluated top-down with the
same floating-point operations (ref: scs.py:555-567,628):

=========  =======================================================================
one        1
depth      parent value + 1 (children of the root: 1)
branch     parent value + (1 if node.length is None else node.length)
bootstrap  node.support (a missing support raises TypeError, as ``None * w`` does there)
=========  =======================================================================
"""

from __future__ import annotations

from collections.abc import Iterable, Sequence
from dataclasses import dataclass

import numpy as np

WEIGHTINGS = ("one", "branch", "depth", "bootstrap")


@dataclass
class LeafTours:
    """Concatenated leaf tours of ``T`` trees over ``n`` vertices (taxa).

    ``adj_depth[i]`` / ``adj_val[i]`` describe the LCA of leaf ``i`` and leaf ``i + 1`` of the
    same tree; the entry of a tree's last leaf is unused (-1 / 0).  ``root_depth[t]`` is the
    depth key that marks "LCA is the root" in tree ``t``.
    """

    n: int
    leaf_offsets: np.ndarray  # int64 [T + 1]
    leaf_taxon: np.ndarray  # int32 [L]
    adj_depth: np.ndarray  # int32 [L]
    adj_val: np.ndarray  # float64 [L]
    root_depth: np.ndarray  # int32 [T]
    tree_weight: np.ndarray  # float64 [T]

    @property
    def num_trees(self) -> int:
        return len(self.tree_weight)

    @property
    def num_leaves(self) -> int:
        return len(self.leaf_taxon)

    def pair_updates(self) -> int:
        """Number of (ordered) leaf-pair visits the row kernel performs: sum of k_t * (k_t - 1)."""
        k = np.diff(self.leaf_offsets)
        return int((k * (k - 1)).sum())


def _node_value(weighting: str, parent_value, node):
    if weighting == "one":
        return 1
    if weighting == "depth":
        return parent_value + 1
    if weighting == "branch":
        return parent_value + (1 if node.length is None else node.length)
    return node.support  # bootstrap


def tour_of_tree(root, taxon_id: dict[str, int], weighting: str):
    """Leaf tour of one tree: (leaf_taxon, adj_depth, adj_val) as Python lists."""
    taxa: list[int] = []
    adj_depth: list[int] = []
    adj_val: list[float] = []
    if not root.children:
        return taxa, adj_depth, adj_val  # a lone tip has no sides (ref: scs.py:570)
    # frame: [node, index of next child, depth, value handed to the children]
    stack = [[root, 0, 0, 0]]
    turn_depth = 0
    turn_val = 0
    while stack:
        frame = stack[-1]
        node, ci = frame[0], frame[1]
        kids = node.children
        if ci == len(kids):
            stack.pop()
            continue
        frame[1] = ci + 1
        if ci:  # coming back to `node` between two of its children: it is the next leaf pair's LCA
            turn_depth, turn_val = frame[2], frame[3]
        child = kids[ci]
        if child.children:
            stack.append([child, 0, frame[2] + 1, _node_value(weighting, frame[3], child)])
        else:
            if taxa:
                if turn_val is None:
                    msg = "unsupported operand type(s) for *: 'NoneType' and 'float'"
                    raise TypeError(msg)
                adj_depth.append(turn_depth)
                adj_val.append(turn_val)
            taxa.append(taxon_id[child.name])
    adj_depth.append(-1)
    adj_val.append(0.0)
    return taxa, adj_depth, adj_val


def flatten_trees(
    trees: Sequence,
    weights: Iterable[float],
    weighting: str,
    taxon_id: dict[str, int],
) -> LeafTours:
    """Leaf tours of ``trees`` with vertex ids taken from ``taxon_id`` (name -> row of W)."""
    if weighting not in WEIGHTINGS:
        msg = f"Invalid weighting strategy selected: '{weighting}'"
        raise ValueError(msg)
    all_taxa: list[int] = []
    all_depth: list[int] = []
    all_val: list[float] = []
    offsets = [0]
    for tree in trees:
        taxa, adj_depth, adj_val = tour_of_tree(tree, taxon_id, weighting)
        all_taxa.extend(taxa)
        all_depth.extend(adj_depth)
        all_val.extend(adj_val)
        offsets.append(len(all_taxa))
    return LeafTours(
        n=len(taxon_id),
        leaf_offsets=np.asarray(offsets, dtype=np.int64),
        leaf_taxon=np.asarray(all_taxa, dtype=np.int32),
        adj_depth=np.asarray(all_depth, dtype=np.int32),
        adj_val=np.asarray(all_val, dtype=np.float64),
        root_depth=np.zeros(len(offsets) - 1, dtype=np.int32),
        tree_weight=np.asarray(list(weights), dtype=np.float64),
    )
