"""GPU: source trees that are not tidy -- internal nodes with a single child (chains of them), polytomies, nodes
without a length, a root with a single child, two-tip trees and lone tips -- against the CPU oracle's whole recursion
(tests/golden/ctrace_untidy_<case>.json.gz, written by tests/golden/make_untidy.py: the reference's own semantics for
such trees, ref: scs.py:560-562, 569-579, 624-631, and ``get_sub_tree``'s merging of unary nodes at every level below
the top, ref: scs.py:444-453).  The fixtures and the synthetic workloads only hold tidy trees; these cases put the
unary-node and missing-value handling of the device-resident forest (csrc/devforest.cu: tours, restriction) and of
its host twin (csrc/forest.cpp) on the same footing."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import compare_with_ctrace, flat_clades, load_ctrace, parse
from spectralclustersupertree_b200.engine import Forest
from spectralclustersupertree_b200.tree import make_tree

pytestmark = pytest.mark.gpu


# case -> nodes where sklearn's k-means returned the same Lloyd-stable but non-optimal split for every seed tried while
# the exact 2-means takes the optimum (reported as mismatches, like s_150x40_one in tests/helpers.py).  The caterpillar
# case is built to have them: a caterpillar's Fiedler coordinate is a chain with many Lloyd-stable cuts (477 of its 916
# spectral nodes have more than one), and at two nodes (m = 433 and m = 39) the reference's restarts never find the best.
EXPECTED_MISMATCHES = {"caterpillar": 2}


@pytest.mark.parametrize("case", ["branch", "bootstrap", "depth", "nocontract", "caterpillar"])
def test_untidy_source_trees_against_the_oracle_trace(engine, case):
    """``nocontract``: the same kind of trees with ``contract_edges=False`` (ref: scs.py:23, 124-133).  ``caterpillar``:
    20 bushy trees + 12 caterpillars of ~750 tips: leaf tours more than 1 024 levels deep, long unary chains to merge."""
    ctrace = load_ctrace(f"untidy_{case}")
    contract = ctrace.get("contract_edges", True)
    trees = parse(ctrace["lines"])
    names = sorted({x for t in trees for x in t.get_tip_names()})
    assert len(names) == ctrace["names"]
    weighting = ctrace["weighting"]

    def build():
        return engine.supertree_build(Forest.from_trees(trees, ctrace["weights"], names), weighting,
                                      contract_edges=contract, record=True)  # fmt: skip

    built = build()
    # every recursion node: components, contracted size, Fiedler eigenvalue (1e-6), bipartition
    report = compare_with_ctrace(built["records"], ctrace, expected_mismatches=EXPECTED_MISMATCHES.get(case, 0))
    divergent = report.pop("divergent_sets")
    assert report["compared"] + report["orphans"] == len(built["records"])
    if not divergent:
        assert len(built["records"]) == len(ctrace["nodes"])
    gid = {name: i for i, name in enumerate(names)}
    reference = {frozenset(gid[x] for x in clade) for clade in make_tree(ctrace["supertree"]).clade_sets()
                 if 1 < len(clade) < len(names)}  # fmt: skip
    ours = {c for c in flat_clades(built["parent"], built["taxon"]) if len(c) < len(names)}
    outside = [c for c in ours ^ reference if not any(c <= d for d in divergent)]
    assert not outside, (len(outside), len(ours ^ reference))
    assert sorted(built["taxon"][built["taxon"] >= 0].tolist()) == list(range(len(names)))
    # the host-forest driver (flat trees restricted by the host threads) takes the same recursion, bit for bit
    engine.set_device_forest(False)
    try:
        on_host = build()
    finally:
        engine.set_device_forest(True)
    assert len(on_host["records"]) == len(built["records"])
    by_taxa = {taxa.tobytes(): (part, stats) for taxa, part, stats in on_host["records"]}
    for taxa, part, stats in built["records"]:
        opart, ostats = by_taxa[taxa.tobytes()]
        assert stats.n_components == ostats.n_components and stats.contracted_size == ostats.contracted_size
        assert np.array_equal(part, opart)
        if stats.n_components == 1 and stats.contracted_size >= 3:
            assert stats.eig[1] == ostats.eig[1], (len(taxa), stats.eig[1], ostats.eig[1])
    assert flat_clades(on_host["parent"], on_host["taxon"]) == flat_clades(built["parent"], built["taxon"])
    print(case, {k: v for k, v in report.items() if k != "divergent_nodes"}, "RF", len(ours ^ reference))


@pytest.mark.parametrize("case", ["branch", "depth"])
def test_untidy_source_trees_through_the_python_recursion(engine, case):
    """The same cases through the reference-shaped loop (``native=False``: one C-ABI call per recursion node, trees
    restricted by ``Forest.induce``): the same supertree as the native driver."""
    from spectralclustersupertree_b200.scs import supertree_of_forest

    ctrace = load_ctrace(f"untidy_{case}")
    trees = parse(ctrace["lines"])
    names = sorted({x for t in trees for x in t.get_tip_names()})
    a = supertree_of_forest(Forest.from_trees(trees, ctrace["weights"], names), ctrace["weighting"], engine=engine, native=True)
    b = supertree_of_forest(Forest.from_trees(trees, ctrace["weights"], names), ctrace["weighting"], engine=engine, native=False)
    assert a.clade_sets() == b.clade_sets()
    assert a.clade_sets() == make_tree(ctrace["supertree"]).clade_sets()


def test_large_untidy_job(engine):
    """5 000 taxa x 300 untidy source trees (branch weighting, tree weights): the top-level node takes the large-node
    path (row kernel over staircases deepened by the unary chains, Lanczos on m ~ 5 000).  W, occurrences and adjacency
    of that node bit for bit against the C oracle, contracted size and Fiedler eigenvalue (1e-6) against the oracle's
    contraction + ARPACK; then the whole job through the device-forest and the host-forest drivers: the same
    recursion, node for node."""
    import sys
    from pathlib import Path

    from oracle import scs_oracle
    from spectralclustersupertree_b200.engine import unpack_bits

    root = Path(__file__).resolve().parent
    for extra in (root / "golden", root.parent / "tools"):
        if str(extra) not in sys.path:
            sys.path.insert(0, str(extra))
    import make_untidy
    from oracle_run import small_eigs

    make_untidy.CASES["large"] = (5000, 300, "branch", 9400, True)
    lines, weights, weighting = make_untidy.untidy_lines("large")
    trees = parse(lines)
    names = sorted({x for t in trees for x in t.get_tip_names()})
    tid = {x: i for i, x in enumerate(names)}
    assert sum(1 for t in trees for node in t.preorder() if len(node.children) == 1) > 10000
    W, C, occ = scs_oracle.pcg_dense_c(trees, weights, weighting, tid)

    forest = Forest.from_trees(trees, weights, names)
    taxa, part, stats = engine.forest_split(forest, weighting, contract_edges=True, seed=1)
    bufs = engine.last_node_buffers()
    assert np.array_equal(bufs["W"], W), "W differs from the oracle"
    assert np.array_equal(bufs["occ"], occ), "occ differs from the oracle"
    assert np.array_equal(unpack_bits(bufs["adj_bits"], len(names)), C > 0), "adjacency differs"
    label = scs_oracle.graph_components(C > 0)
    assert stats.n_components == len(np.unique(label))
    if stats.n_components == 1:
        _, Wc, _ = scs_oracle.contract_dense(W, C, occ)
        assert stats.contracted_size == Wc.shape[0]
        lam2, _, _ = small_eigs(Wc)
        assert abs(stats.eig[1] - lam2) < 1e-6, (stats.eig[1], lam2)
    top_m = int(stats.contracted_size)
    del W, C, bufs

    def build():
        return engine.supertree_build(Forest.from_trees(trees, weights, names), weighting, record=True)

    on_device = build()
    engine.set_device_forest(False)
    try:
        on_host = build()
    finally:
        engine.set_device_forest(True)
    assert len(on_device["records"]) == len(on_host["records"])
    by_taxa = {taxa.tobytes(): (part, stats) for taxa, part, stats in on_host["records"]}
    for taxa, part, stats in on_device["records"]:
        opart, ostats = by_taxa[taxa.tobytes()]
        assert stats.n_components == ostats.n_components and stats.contracted_size == ostats.contracted_size
        assert np.array_equal(part, opart), len(taxa)
        if stats.n_components == 1 and stats.contracted_size >= 3:
            assert stats.eig[1] == ostats.eig[1], (len(taxa), stats.eig[1], ostats.eig[1])
    assert flat_clades(on_device["parent"], on_device["taxon"]) == flat_clades(on_host["parent"], on_host["taxon"])
    assert sorted(on_device["taxon"][on_device["taxon"] >= 0].tolist()) == list(range(len(names)))
    print("large untidy:", len(on_device["records"]), "recursion nodes, top-level m", top_m)
