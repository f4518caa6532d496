"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's hot path on dense arrays.

This is the oracle the CUDA path is checked against (tests/, smoke()) and the CPU baseline
bench.py times.  It is never imported by the product package.

Each function restates one reference function (``ref:`` = /root/reference/src/sc_supertree/scs.py)
with the dict-of-tuples graph replaced by dense arrays indexed by *vertex id = rank of the taxon
name in sorted(all names)*:

====================================  ==========================================
``pcg_dense``                         ref: scs.py:495-583 + 586-663
``graph_components``                  ref: scs.py:458-492
``contract_dense``                    ref: scs.py:261-387
``spectral_bipartition``              ref: scs.py:210-258 (sklearn call restated verbatim)
``construct_supertree``               ref: scs.py:18-174 (recursion + glue :177-207,:390-455,:708-746)
====================================  ==========================================

The spectral stage lives in third-party code that is installed in this image and on the GPU box
(scikit-learn 1.9.0 -> scipy 1.18.1 ARPACK), so the oracle calls
``sklearn.cluster.SpectralClustering`` exactly as the reference does rather than restating ARPACK.
``normalized_affinity_eigs`` adds an independent dense ``eigh`` of the same operator for the
eigenvalue tolerance check.

Pinning: ``tests/golden/make_golden.py`` records the outputs of the reference's own functions
(imported unmodified through ``oracle/cogent3_shim.py`` in the build container) and
``tests/test_oracle_golden.py`` checks every function here against those vectors, and against
the reference's golden ``.tre`` files.  ``pcg_dense`` has a C twin in ``oracle/pcg_oracle.c``
(same algorithm, same summation order) used when the Python loops would be too slow.
"""

from __future__ import annotations

import ctypes
from collections.abc import Sequence
from pathlib import Path

import numpy as np

from spectralclustersupertree_b200.tree import NotCompleted, PhyloNode

WEIGHTINGS = ("one", "branch", "depth", "bootstrap")


# ---------------------------------------------------------------------------------------------
# proper cluster graph (ref: scs.py:495-663)
# ---------------------------------------------------------------------------------------------
def _length_function(weighting: str):
    # ref: scs.py:555-567
    if weighting == "one":
        return lambda _length, _node: 1
    if weighting == "depth":
        return lambda length, _node: length + 1
    if weighting == "branch":
        return lambda length, node: length + (1 if node.length is None else node.length)
    if weighting == "bootstrap":
        return lambda _length, node: node.support
    msg = f"Invalid weighting strategy selected: '{weighting}'"
    raise ValueError(msg)


def pcg_dense(trees: Sequence[PhyloNode], weights: Sequence[float], weighting: str, taxon_id: dict[str, int]):
    """Dense proper-cluster-graph: (W float64 n x n, C int32 n x n, occ int32 n).

    ``W[a, b]`` accumulates ``length * tree_weight`` in tree input order for every tree in which
    a and b sit under the same child of the root, ``length`` being the value the reference's
    ``length_function`` gives their LCA (ref: scs.py:569-581, 628, 644-658); ``C`` counts those
    trees; ``occ[a]`` counts the trees containing ``a`` below one of the root's children.
    """
    f = _length_function(weighting)
    n = len(taxon_id)
    W = np.zeros((n, n), dtype=np.float64)
    C = np.zeros((n, n), dtype=np.int32)
    occ = np.zeros(n, dtype=np.int32)
    for tree, weight in zip(trees, weights, strict=True):
        for side in tree:  # ref: scs.py:570
            side_taxa = _dfs_pairs(W, C, side, weight, 0, f, taxon_id)
            occ[side_taxa] += 1  # ref: scs.py:580-581
    return W, C, occ


def _dfs_pairs(W, C, top, tree_weight, length, f, taxon_id) -> np.ndarray:
    """ref: scs.py:586-663 without recursion: returns the vertex ids of the tips below ``top``."""
    # first pass: top-down values (ref: scs.py:628); second pass: bottom-up tip lists + pair updates
    order: list[tuple[PhyloNode, float]] = []
    stack = [(top, length)]
    while stack:
        node, above = stack.pop()
        if node.is_tip():
            order.append((node, above))
            continue
        here = f(above, node)
        order.append((node, here))
        stack.extend((child, here) for child in node)
    tips_below: dict[int, np.ndarray] = {}
    for node, here in reversed(order):
        if node.is_tip():
            tips_below[id(node)] = np.array([taxon_id[node.name]], dtype=np.int64)
            continue
        children_tips = [tips_below.pop(id(child)) for child in node]
        for i in range(1, len(children_tips)):  # ref: scs.py:644-658
            for j in range(i):
                rows, cols = children_tips[i], children_tips[j]
                term = here * tree_weight  # one rounding, then one add: never fused
                W[np.ix_(rows, cols)] += term
                W[np.ix_(cols, rows)] += term
                C[np.ix_(rows, cols)] += 1
                C[np.ix_(cols, rows)] += 1
        tips_below[id(node)] = np.concatenate(children_tips)
    return tips_below[id(top)]


_C_LIB = None


def _c_lib():
    """The C twin of ``pcg_dense`` (oracle/pcg_oracle.c), built by ``oracle/build.py``."""
    global _C_LIB  # noqa: PLW0603
    if _C_LIB is None:
        path = Path(__file__).parent / "_build" / "libpcg_oracle.so"
        if not path.is_file():
            from oracle import build as _build

            _build.build()
        lib = ctypes.CDLL(str(path))
        lib.pcg_oracle_dense.restype = ctypes.c_int
        lib.pcg_oracle_rows.restype = ctypes.c_int
        _C_LIB = lib
    return _C_LIB


def children_csr(trees: Sequence[PhyloNode], taxon_id: dict[str, int], weighting: str):
    """Trees as child lists for the C oracle: independent of the product's leaf-tour flattening."""
    child_ptr = [0]
    child_idx: list[int] = []
    tip_taxon: list[int] = []
    own: list[float] = []  # the per-node quantity length_function reads
    roots: list[int] = []
    for tree in trees:
        index: dict[int, int] = {}
        nodes = list(tree.preorder())
        base = len(tip_taxon)
        for k, node in enumerate(nodes):
            index[id(node)] = base + k
        roots.append(base)
        for node in nodes:
            tip_taxon.append(taxon_id[node.name] if node.is_tip() else -1)
            if weighting == "branch":
                own.append(1.0 if node.length is None else float(node.length))
            elif weighting == "bootstrap":
                if node.support is None:
                    own.append(float("nan"))
                else:
                    own.append(float(node.support))
            else:
                own.append(0.0)
            child_idx.extend(index[id(c)] for c in node)
            child_ptr.append(len(child_idx))
    return (
        np.asarray(roots, dtype=np.int64),
        np.asarray(child_ptr, dtype=np.int64),
        np.asarray(child_idx, dtype=np.int64),
        np.asarray(tip_taxon, dtype=np.int32),
        np.asarray(own, dtype=np.float64),
    )


def pcg_dense_c(trees, weights, weighting: str, taxon_id: dict[str, int]):
    """Same result as ``pcg_dense`` (bit for bit), computed by oracle/pcg_oracle.c."""
    roots, child_ptr, child_idx, tip_taxon, own = children_csr(trees, taxon_id, weighting)
    return pcg_dense_c_arrays(len(taxon_id), roots, child_ptr, child_idx, tip_taxon, own, weights, weighting)


def pcg_dense_c_arrays(n, roots, child_ptr, child_idx, tip_taxon, own, weights, weighting: str):
    lib = _c_lib()
    W = np.zeros((n, n), dtype=np.float64)
    C = np.zeros((n, n), dtype=np.int32)
    occ = np.zeros(n, dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(list(weights), dtype=np.float64))
    mode = WEIGHTINGS.index(weighting)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.pcg_oracle_dense(
        ctypes.c_int64(n), ctypes.c_int64(len(roots)), p(roots), p(child_ptr), p(child_idx), p(tip_taxon),
        p(own), p(w), ctypes.c_int(mode), p(W), p(C), p(occ),
    )  # fmt: skip
    if rc != 0:
        msg = f"pcg_oracle_dense failed with code {rc}"
        raise RuntimeError(msg)
    return W, C, occ


def pcg_rows_c_arrays(n, roots, child_ptr, child_idx, tip_taxon, own, weights, weighting: str, row_lo: int, row_hi: int):
    """Rows [row_lo, row_hi) of the dense W and C (oracle/pcg_oracle.c:pcg_oracle_rows): the same ordered sums as
    ``pcg_dense_c_arrays`` without the n x n memory, for checks at 50 000 taxa."""
    lib = _c_lib()
    rows = row_hi - row_lo
    W = np.zeros((rows, n), dtype=np.float64)
    C = np.zeros((rows, n), dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(list(weights), dtype=np.float64))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.pcg_oracle_rows(
        ctypes.c_int64(n), ctypes.c_int64(len(roots)), p(roots), p(child_ptr), p(child_idx), p(tip_taxon), p(own),
        p(w), ctypes.c_int(WEIGHTINGS.index(weighting)), ctypes.c_int64(row_lo), ctypes.c_int64(row_hi), p(W), p(C),
    )  # fmt: skip
    if rc != 0:
        msg = f"pcg_oracle_rows failed with code {rc}"
        raise RuntimeError(msg)
    return W, C


# ---------------------------------------------------------------------------------------------
# components (ref: scs.py:458-492)
# ---------------------------------------------------------------------------------------------
def graph_components(adjacency: np.ndarray) -> np.ndarray:
    """Component label of every vertex = smallest vertex id in its component.

    ``adjacency`` is a boolean n x n matrix; the reference explores the ``edges`` sets, i.e.
    "co-occurred at least once" (``C > 0``), not ``W > 0`` (ref: scs.py:479-490, 651-652).
    """
    n = adjacency.shape[0]
    if n >= 256:
        # same labels as the depth-first search below, from scipy's C implementation (the search itself, done in
        # Python as the reference does it, costs 8 s at n = 10 000 -- a CPU baseline should not be charged for that)
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import connected_components

        _, comp = connected_components(csr_matrix(adjacency), directed=False)
        smallest = np.full(comp.max() + 1, n, dtype=np.int64)
        np.minimum.at(smallest, comp, np.arange(n))
        return smallest[comp].astype(np.int32)
    label = np.full(n, -1, dtype=np.int32)
    for start in range(n):
        if label[start] >= 0:
            continue
        label[start] = start
        frontier = [start]
        while frontier:
            current = frontier.pop()
            for neighbour in np.flatnonzero(adjacency[current]):
                if label[neighbour] < 0:
                    label[neighbour] = start
                    frontier.append(int(neighbour))
    return label


# ---------------------------------------------------------------------------------------------
# contraction (ref: scs.py:261-387)
# ---------------------------------------------------------------------------------------------
def contract_dense(W: np.ndarray, C: np.ndarray, occ: np.ndarray):
    """Contract taxa that always appear together.

    Returns ``(group, Wc, Ac)``: ``group[v]`` = id of the contracted vertex holding ``v``,
    numbered by smallest member; ``Wc[A, B]`` = max of ``W[u, v]`` over the existing edges
    between the groups (ref: scs.py:382-387), 0 where there is none; ``Ac`` the contracted
    adjacency.  The max-graph keeps pair (u, v) iff ``C[u, v] == max(occ[u], occ[v])`` and the
    pair co-occurred at all (ref: scs.py:302-313 iterates ``taxa_co_occurrences`` only).
    """
    n = len(occ)
    exists = C > 0
    top = np.maximum(occ[:, None], occ[None, :])
    keep = exists & (top == C)
    rep = graph_components(keep)
    _, group = np.unique(rep, return_inverse=True)  # ids ordered by smallest member
    group = group.astype(np.int32)
    m = int(group.max()) + 1 if n else 0
    Wc = np.zeros((m, m), dtype=np.float64)
    Ac = np.zeros((m, m), dtype=bool)
    best = np.full((m, m), -np.inf)
    us, vs = np.nonzero(exists)
    gu, gv = group[us], group[vs]
    outside = gu != gv  # edges inside a contraction vanish (ref: scs.py:352-354)
    np.maximum.at(best, (gu[outside], gv[outside]), W[us[outside], vs[outside]])
    Ac = np.isfinite(best)
    Wc[Ac] = best[Ac]
    return group, Wc, Ac


# ---------------------------------------------------------------------------------------------
# spectral bipartition (ref: scs.py:210-258)
# ---------------------------------------------------------------------------------------------
def spectral_bipartition(W: np.ndarray, random_state: np.random.RandomState) -> np.ndarray:
    """Labels in {0, 1} from sklearn, called exactly as the reference calls it (ref: scs.py:235-252)."""
    from sklearn.cluster import SpectralClustering

    sc = SpectralClustering(
        2,
        affinity="precomputed",
        assign_labels="kmeans",
        n_jobs=1,
        random_state=random_state,
    )
    return np.asarray(sc.fit_predict(np.array(W, dtype=np.float64)), dtype=np.int32)


def normalized_affinity_eigs(W: np.ndarray, k: int = 3) -> tuple[np.ndarray, np.ndarray]:
    """The k smallest eigenpairs of L = I - D^-1/2 W D^-1/2 by dense ``eigh``.

    Same operator as scipy's ``csgraph.laplacian(normed=True)`` builds for sklearn (zero
    diagonal, degree = column sum, isolated vertices scaled by 1; scipy _laplacian.py:532-563).
    Returns (eigenvalues ascending, embedding columns ``x_k / sqrt(d)``).
    """
    from scipy.linalg import eigh

    A = np.array(W, dtype=np.float64)
    np.fill_diagonal(A, 0.0)
    d = A.sum(axis=0)
    s = np.where(d == 0, 1.0, np.sqrt(d))
    L = -A / s[:, None] / s[None, :]
    np.fill_diagonal(L, 1.0)
    k = min(k, len(d))
    vals, vecs = eigh(L, subset_by_index=(0, k - 1))
    return vals, vecs / s[:, None]


# ---------------------------------------------------------------------------------------------
# the recursion (ref: scs.py:18-174) on dense node graphs
# ---------------------------------------------------------------------------------------------
def construct_supertree(
    trees: Sequence[PhyloNode],
    weights: Sequence[float] | None = None,
    pcg_weighting: str = "one",
    *,
    contract_edges: bool = True,
    random_state: np.random.RandomState | None = None,
    use_c: bool = False,
    trace: list | None = None,
    timers: dict | None = None,
    node_hook=None,
    defer: list | None = None,
    defer_max_taxa: int = 0,
) -> PhyloNode:
    """CPU restatement of the reference's ``construct_supertree`` (ref: scs.py:18-174).

    ``trace``, when given, receives one dict per recursion node (sorted vertex names, number of
    components, spectral partition) for node-by-node parity checks.  ``timers`` accumulates wall-clock
    seconds per stage (``pcg``, ``components``, ``contract``, ``spectral``, ``induce``, ``hook``);
    ``node_hook(record, Wc, side, groups)`` is called at every spectral node (its time goes to
    ``hook`` and is therefore separable from the reference-equivalent work); if it returns an array, the
    recursion continues with those labels instead of sklearn's.  ``defer`` (a list) makes the recursion stop
    at components of at most ``defer_max_taxa`` taxa: a placeholder tip is put where the sub-problem's supertree
    belongs and ``(placeholder, trees, weights)`` is appended (``construct_supertree_parallel`` solves them in a
    process pool).
    """
    import time as _time

    def _lap(key, since):
        if timers is not None:
            timers[key] = timers.get(key, 0.0) + (_time.perf_counter() - since)
        return _time.perf_counter()

    if random_state is None:
        random_state = np.random.RandomState()
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
    pairs = [(t, w) for t, w in zip(trees, weights, strict=True) if not isinstance(t, NotCompleted)]
    if len(pairs) == 0:
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    trees, weights = zip(*pairs, strict=True)

    if len(trees) == 1:  # ref: scs.py:96-98
        from spectralclustersupertree_b200.tree import make_tree

        for node in trees[0].iter_nontips(include_self=True):
            node.name = ""
        return make_tree(trees[0].get_newick())

    all_names: set[str] = set()
    for tree in trees:
        all_names.update(tree.get_tip_names())
    if len(all_names) <= 2:
        return _star(all_names)

    names = sorted(all_names)
    taxon_id = {name: i for i, name in enumerate(names)}
    build = pcg_dense_c if use_c else pcg_dense
    mark = _time.perf_counter()
    W, C, occ = build(trees, weights, pcg_weighting, taxon_id)
    mark = _lap("pcg", mark)
    label = graph_components(C > 0)
    reps = np.unique(label)
    mark = _lap("components", mark)
    record = {"names": names, "n_components": len(reps)}
    if len(reps) == 1:
        groups = np.arange(len(names), dtype=np.int32)
        Wc = W
        if contract_edges:
            groups, Wc, _ = contract_dense(W, C, occ)
        mark = _lap("contract", mark)
        side = spectral_bipartition(Wc, random_state)
        mark = _lap("spectral", mark)
        record["contracted_size"] = int(Wc.shape[0])
        if node_hook is not None:
            # the hook may return replacement labels (tools/oracle_run.py steers RNG-dependent k-means nodes)
            replacement = node_hook(record, Wc, side, groups)
            if replacement is not None:
                side = replacement
            mark = _lap("hook", mark)
        parts = [{names[v] for v in range(len(names)) if side[groups[v]] == s} for s in (0, 1)]
        record["partition"] = [sorted(p) for p in parts]
    else:
        parts = [{names[v] for v in np.flatnonzero(label == r)} for r in reps]
    if trace is not None:
        trace.append(record)
    del W, C

    child_trees: list[PhyloNode] = []
    for component in parts:
        if len(component) <= 2:
            child_trees.append(_star(component))
            continue
        new_trees, new_weights = [], []
        mark = _time.perf_counter()
        for tree, weight in zip(trees, weights, strict=True):  # ref: scs.py:444-453
            if len(component.intersection(tree.get_tip_names())) < 2:
                continue
            sub = tree.get_sub_tree(component, ignore_missing=True, as_rooted=True)
            sub.name = "root"
            new_trees.append(sub)
            new_weights.append(weight)
        _lap("induce", mark)
        if defer is not None and len(component) <= defer_max_taxa and len(new_trees) > 0:
            placeholder = PhyloNode(f"__deferred_{len(defer)}__")
            defer.append((placeholder, new_trees, new_weights))
            child_trees.append(placeholder)
        else:
            child_trees.append(
                construct_supertree(
                    new_trees,
                    new_weights,
                    pcg_weighting,
                    contract_edges=contract_edges,
                    random_state=random_state,
                    use_c=use_c,
                    trace=trace,
                    timers=timers,
                    node_hook=node_hook,
                    defer=defer,
                    defer_max_taxa=defer_max_taxa,
                )
            )
        seen: set[str] = set()
        for tree in new_trees:
            seen.update(tree.get_tip_names())
        child_trees.extend(_star((x,)) for x in sorted(component.difference(seen)))
    return _connect(child_trees)


# ---- the same recursion with the independent sub-problems spread over the host cores ------------------
# The reference is single-process (ref: scs.py:239 n_jobs=1); sub-problems below a split are independent
# (ref: scs.py:139-166), so a CPU baseline "with all the host threads it can use" runs the top of the recursion
# serially (its BLAS calls use every core) and hands the components of at most ``defer_max_taxa`` taxa to a pool
# of forked worker processes.  Used by bench.py's reference arm only.
_DEFERRED: list = []
_DEFERRED_ARGS: dict = {}


def _solve_deferred(index: int):
    import time as _time

    _placeholder, trees, weights = _DEFERRED[index]
    timers: dict = {}
    trace: list = []
    t0 = _time.perf_counter()
    try:
        from threadpoolctl import threadpool_limits

        limiter = threadpool_limits(limits=1)  # one BLAS thread per worker process
    except Exception:  # noqa: BLE001
        limiter = None
    tree = construct_supertree(
        trees, weights, _DEFERRED_ARGS["weighting"], contract_edges=_DEFERRED_ARGS["contract_edges"],
        random_state=np.random.RandomState(1 + index), use_c=True, trace=trace, timers=timers,
    )  # fmt: skip
    if limiter is not None:
        limiter.restore_original_limits()
    spectral = sum(1 for r in trace if "partition" in r)
    return index, tree.get_newick(), len(trace), spectral, timers, _time.perf_counter() - t0


def construct_supertree_parallel(trees, weights, pcg_weighting: str, *, contract_edges: bool = True,
                                 workers: int = 0, defer_max_taxa: int = 1500, timers: dict | None = None,
                                 info: dict | None = None) -> PhyloNode:  # fmt: skip
    """``construct_supertree`` with sub-problems of at most ``defer_max_taxa`` taxa solved by ``workers`` forked
    processes (0 = one per host core).  Same algorithm per node; the RNG stream of a sub-problem is seeded by
    its index instead of continuing the parent's."""
    import multiprocessing as mp
    import os

    from spectralclustersupertree_b200.tree import make_tree

    global _DEFERRED, _DEFERRED_ARGS  # noqa: PLW0603
    workers = workers or (os.cpu_count() or 1)
    deferred: list = []
    trace: list = []
    top = construct_supertree(
        trees, weights, pcg_weighting, contract_edges=contract_edges, random_state=np.random.RandomState(0),
        use_c=True, trace=trace, timers=timers, defer=deferred, defer_max_taxa=defer_max_taxa,
    )  # fmt: skip
    nodes, spectral = len(trace), sum(1 for r in trace if "partition" in r)
    pool_timers: dict = {}
    if deferred:
        _DEFERRED = deferred
        _DEFERRED_ARGS = {"weighting": pcg_weighting, "contract_edges": contract_edges}
        order = sorted(range(len(deferred)), key=lambda i: -sum(len(t.get_tip_names()) for t in deferred[i][1]))
        with mp.get_context("fork").Pool(min(workers, len(deferred))) as pool:
            for index, newick, n_nodes, n_spectral, part_timers, _seconds in pool.imap_unordered(_solve_deferred, order):
                sub = make_tree(newick)
                placeholder = deferred[index][0]
                placeholder.name = sub.name
                placeholder.children = []
                for child in list(sub.children):
                    placeholder.append(child)
                nodes += n_nodes
                spectral += n_spectral
                for key, value in part_timers.items():
                    pool_timers[key] = pool_timers.get(key, 0.0) + value
        _DEFERRED = []
    if info is not None:
        info.update({"recursion_nodes": nodes, "spectral_nodes": spectral, "deferred_subproblems": len(deferred),
                     "workers": workers, "pool_cpu_seconds": pool_timers})  # fmt: skip
    return top


def _star(names) -> PhyloNode:
    # ref: scs.py:728-746
    return _connect([PhyloNode(name) for name in sorted(names)])


def _connect(trees: Sequence[PhyloNode]) -> PhyloNode:
    # ref: scs.py:390-408
    if len(trees) == 1:
        return trees[0]
    return PhyloNode("root", trees)


def rf_distance(a: PhyloNode, b: PhyloNode) -> int:
    """Rooted Robinson-Foulds distance: size of the symmetric difference of the clade sets."""
    return len(a.clade_sets() ^ b.clade_sets())
