"""Drop-in ``construct_supertree`` whose per-node hot path runs on the GPU.

Mirrors the reference's public function (ref: src/sc_supertree/scs.py:18-174): same signature,
defaults, error messages and tree-shaped result.  The divide-and-conquer over taxon sets stays on
the host (an explicit stack instead of Python recursion, so deep supertrees do not hit the
recursion limit); at every recursion node the four private functions the reference calls
(ref: scs.py:110-134) are one call into ``libscs_b200.so``:

====================================  ==========================================================
reference (scs.py)                    here
====================================  ==========================================================
``_proper_cluster_graph_edges``       CUDA row kernel over leaf tours        (csrc/pcg.cu)
``_get_graph_components``             CUDA union-find on the adjacency bits  (csrc/components.cu)
``_contract_proper_cluster_graph``    CUDA max-merge                         (csrc/contract.cu)
``spectral_cluster_graph``            CUDA Lanczos + exact 2-means           (csrc/spectral.cu)
``_generate_induced_trees_…``         flat-array restriction on the device   (csrc/devforest.cu;
                                      host twin for ``native=False``: csrc/forest.cpp)
====================================  ==========================================================

There is no CPU fallback: without ``libscs_b200.so`` or a CUDA device the call raises.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Literal

import numpy as np

from .engine import Engine, Forest, default_engine
from .load import untouched_forest
from .tree import NotCompleted, PhyloNode, make_tree

WEIGHTINGS = ("one", "branch", "depth", "bootstrap")


def _is_not_completed(obj) -> bool:
    if isinstance(obj, NotCompleted):
        return True
    return type(obj).__name__ == "NotCompleted"  # a real cogent3 NotCompleted, when cogent3 is installed


def _in_the_callers_class(result: PhyloNode, trees: Sequence) -> PhyloNode:
    """The supertree as a cogent3 tree when the source trees are cogent3 trees (where cogent3 is installed the
    reference returns its ``PhyloNode``, ref: scs.py:25,390-408); the package's own node class otherwise."""
    module = type(trees[0]).__module__ or ""
    if module != "cogent3" and not module.startswith("cogent3."):
        return result
    try:
        import cogent3
    except ImportError:
        return result
    if getattr(cogent3, "_scs_b200_shim", False):  # oracle/cogent3_shim.py: the stand-in, not the library
        return result
    return cogent3.make_tree(result.get_newick())


def construct_supertree(
    trees: Sequence[PhyloNode],
    weights: Sequence[float] | None = None,
    pcg_weighting: Literal["one", "branch", "depth", "bootstrap"] = "one",
    *,
    contract_edges: bool = True,
    random_state: np.random.RandomState | None = None,
    engine: Engine | None = None,
    trace: list | None = None,
) -> PhyloNode:
    """Spectral Cluster Supertree (SCS) of ``trees`` (ref: scs.py:18-59).

    Parameters are the reference's.  ``random_state`` only seeds the eigensolver's start vector:
    the GPU bipartition is deterministic for a given seed (default seed 0 when None).
    Two keyword-only extras, not in the reference: ``engine`` selects the GPU context and
    ``trace`` (a list) receives one record per recursion node for parity checks.
    """
    if len(trees) == 0:
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    if pcg_weighting not in WEIGHTINGS:
        msg = f"Invalid weighting strategy selected: '{pcg_weighting}'"
        raise ValueError(msg)
    if weights is None:
        weights = [1.0 for _ in range(len(trees))]
    if len(trees) != len(weights):
        msg = f"The number of trees ({len(trees)}) and tree weights ({len(weights)}) must match."
        raise ValueError(msg)
    # the list one load_trees call returned, untouched: its file was parsed natively into the flat store already
    # (load.LoadedTree); building and flattening a million node objects would only reproduce that forest
    loaded = untouched_forest(trees, weights) if len(trees) > 1 else None
    if loaded is not None:
        if len(loaded.names) <= 2:  # ref: scs.py:105-106
            return _star(loaded.names)
        seed = 0 if random_state is None else int(random_state.randint(0, 2**31 - 1))
        return supertree_of_forest(
            loaded, pcg_weighting, contract_edges=contract_edges, seed=seed, engine=engine, trace=trace
        )
    pairs = [(t, w) for t, w in zip(trees, weights, strict=True) if not _is_not_completed(t)]
    if len(pairs) == 0:
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    trees, weights = zip(*pairs, strict=True)

    if len(trees) == 1:  # ref: scs.py:96-98
        for node in trees[0].iter_nontips(include_self=True):
            node.name = ""
        return _in_the_callers_class(make_tree(trees[0].get_newick()), trees)

    # one walk over the node objects yields the flat forest and the taxon names (ref: scs.py:100-103 collects the
    # names with a get_tip_names() pass of its own)
    forest = Forest.from_trees(trees, [float(w) for w in weights])
    if len(forest.names) <= 2:  # ref: scs.py:105-106
        return _in_the_callers_class(_star(forest.names), trees)

    seed = 0 if random_state is None else int(random_state.randint(0, 2**31 - 1))
    result = supertree_of_forest(
        forest, pcg_weighting, contract_edges=contract_edges, seed=seed, engine=engine, trace=trace
    )
    return _in_the_callers_class(result, trees)


def supertree_of_forest(
    forest: Forest,
    pcg_weighting: str = "one",
    *,
    contract_edges: bool = True,
    seed: int = 0,
    engine: Engine | None = None,
    trace: list | None = None,
    node_hook=None,
    native: bool | None = None,
) -> PhyloNode:
    """The recursion of ``construct_supertree`` (ref: scs.py:96-174) on a flat ``Forest``.

    By default the recursion itself runs natively (``csrc/driver.cu``): breadth-first over the
    independent sub-problems, all small frontier nodes in one batched launch.  ``native=False`` runs
    the same recursion from Python, one C-ABI call per node (the reference-shaped loop below);
    ``node_hook(forest, seed)``, if given, is called for every recursion node that reaches the GPU,
    just before it is split (bench.py records the nodes with it) and implies ``native=False``."""
    if pcg_weighting not in WEIGHTINGS:
        msg = f"Invalid weighting strategy selected: '{pcg_weighting}'"
        raise ValueError(msg)
    if engine is None:
        engine = default_engine()
    names = forest.names
    if native is None:
        native = node_hook is None
    if native:
        built = engine.supertree_build(forest, pcg_weighting, contract_edges=contract_edges, seed=seed,
                                       record=trace is not None)  # fmt: skip
        if trace is not None:
            for taxa, part, stats in built["records"]:
                record = {"names": [names[x] for x in taxa], "n_components": int(stats.n_components)}
                if stats.n_components == 1:
                    record["contracted_size"] = int(stats.contracted_size)
                    record["partition"] = [[names[x] for x in taxa[part == c]] for c in (0, 1)]
                    record["stats"] = stats.as_dict()
                trace.append(record)
        return _tree_from_flat(built["parent"], built["taxon"], names)
    holder = PhyloNode("holder")
    holder.children = [None]
    stack: list[tuple[Forest, PhyloNode, int]] = [(forest, holder, 0)]
    node_counter = 0

    def place(parent: PhyloNode, slot: int, child: PhyloNode) -> None:
        child.parent = parent
        parent.children[slot] = child

    while stack:
        current, parent, slot = stack.pop()
        if current.num_trees == 0:  # ref: scs.py:63-65 reached through the recursion
            msg = "There must be at least one tree to make a supertree."
            raise ValueError(msg)
        if current.num_trees == 1:  # ref: scs.py:96-98
            place(parent, slot, _tree_from_arrays(*current.tree_arrays(0), names))
            continue
        present = current.taxa()
        if len(present) <= 2:  # ref: scs.py:105-106
            place(parent, slot, _star([names[x] for x in present]))
            continue
        if node_hook is not None:
            node_hook(current, seed + node_counter)
        taxa, part, stats = engine.forest_split(
            current, pcg_weighting, contract_edges=contract_edges, seed=seed + node_counter
        )
        node_counter += 1
        n_parts = stats.n_components if stats.n_components != 1 else 2
        order = np.argsort(part, kind="stable")
        bounds = np.searchsorted(part[order], np.arange(n_parts + 1))
        components = [taxa[order[bounds[c] : bounds[c + 1]]] for c in range(n_parts)]
        if trace is not None:
            record = {"names": [names[x] for x in taxa], "n_components": int(stats.n_components)}
            if stats.n_components == 1:
                record["contracted_size"] = int(stats.contracted_size)
                record["partition"] = [[names[x] for x in comp] for comp in components]
                record["stats"] = stats.as_dict()
            trace.append(record)

        # ref: scs.py:136-171 -- children of this node, in component order
        plan: list[tuple[str, object]] = []
        for comp in components:
            if len(comp) <= 2:  # ref: scs.py:143-145
                plan.append(("tree", _star([names[x] for x in comp])))
                continue
            child = current.induce(comp)
            plan.append(("forest", child))
            if child.num_trees > 0:
                covered = child.taxa()
                if len(covered) != len(comp):  # ref: scs.py:168-171
                    plan.extend(("tree", PhyloNode(names[x])) for x in np.setdiff1d(comp, covered))
            # an empty child forest raises when popped, exactly where the reference's recursion would
        node = PhyloNode("root")
        node.children = [None] * len(plan)
        place(parent, slot, node)
        for i, (kind, item) in enumerate(plan):
            if kind == "tree":
                place(node, i, item)
            else:
                stack.append((item, node, i))
    result = holder.children[0]
    result.parent = None
    return result


def flat_newick(parent, taxon, names: Sequence[str]) -> str:
    """Newick text of the native driver's flat result, written natively (``scs_flat_tree_newick``): what
    ``_tree_from_flat(...).get_newick()`` gives, without a node object per node."""
    import ctypes

    from . import _lib
    from ._lib import ptr

    lib = _lib.load()
    blob = b"".join(name.encode("utf-8") + b"\0" for name in names)
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    taxon = np.ascontiguousarray(taxon, dtype=np.int32)
    text, size = ctypes.c_void_p(), ctypes.c_size_t()
    status = lib.scs_flat_tree_newick(len(parent), ptr(parent), ptr(taxon), blob, len(blob), len(names),
                                      ctypes.byref(text), ctypes.byref(size))  # fmt: skip
    if status != 0:
        msg = f"scs_flat_tree_newick failed with status {status}"
        raise RuntimeError(msg)
    try:
        return ctypes.string_at(text.value, size.value).decode("utf-8")
    finally:
        lib.scs_free(text)


def supertree_newick_of_forest(forest: Forest, pcg_weighting: str = "one", *, contract_edges: bool = True, seed: int = 0,
                               engine: Engine | None = None) -> str:
    """``supertree_of_forest`` for callers that only want the Newick text (the ``scs`` command): flat forest in,
    native recursion, native Newick emission -- no node objects at either end."""
    if pcg_weighting not in WEIGHTINGS:
        msg = f"Invalid weighting strategy selected: '{pcg_weighting}'"
        raise ValueError(msg)
    if engine is None:
        engine = default_engine()
    built = engine.supertree_build(forest, pcg_weighting, contract_edges=contract_edges, seed=seed)
    return flat_newick(built["parent"], built["taxon"], forest.names)


def _tree_from_flat(parent, taxon, names: Sequence[str]) -> PhyloNode:
    """PhyloNode tree of the native driver's flat result (parent[i] < i, children in index order)."""
    nodes = [PhyloNode(names[x]) if x >= 0 else PhyloNode("root") for x in taxon.tolist()]
    for k, up in enumerate(parent.tolist()):
        if up >= 0:
            nodes[up].append(nodes[k])
    return nodes[0]


def _star(names: Sequence[str]) -> PhyloNode:
    """ref: scs.py:728-746 followed by ``_connect_trees`` (:390-408): one tip stays a tip."""
    tips = [PhyloNode(name) for name in names]
    if len(tips) == 1:
        return tips[0]
    return PhyloNode("root", tips)


def _tree_from_arrays(parent, length, support, taxon, names: Sequence[str]) -> PhyloNode:  # noqa: ARG001
    """Topology-only copy of a flat tree: the single-tree shortcut (ref: scs.py:96-98, 177-187)."""
    nodes = [PhyloNode(names[x] if x >= 0 else "") for x in taxon]
    for k in range(1, len(nodes)):
        nodes[parent[k]].append(nodes[k])
    root = nodes[0]
    counter = 0
    for node in nodes:  # make_tree names what the Newick round trip left blank
        if node.children:
            if node is root:
                node.name = "root"
            else:
                node.name = f"edge.{counter}"
                counter += 1
    return root
