"""GPU: the drop-in ``construct_supertree`` / ``load_trees`` / ``scs`` against the reference's
known-answer tests (ref: tests/test_spectral_cluster_supertree.py, tests/test_cli.py), its fixture
files, and node-by-node against traces of the reference's own runs."""

from __future__ import annotations

import json

import numpy as np
import pytest

from helpers import GOLDEN, compare_with_reference_trace, kat_cases, load_case, parse, rf, rf_outside
from spectralclustersupertree_b200 import construct_supertree, load_trees
from spectralclustersupertree_b200.tree import NotCompleted, make_tree

pytestmark = pytest.mark.gpu


def scs_test(in_trees, expected, engine, **kwargs):
    """ref: tests/test_spectral_cluster_supertree.py:12-27"""
    result = construct_supertree([make_tree(s) for s in in_trees], engine=engine, **kwargs).sorted()
    assert result.same_shape(make_tree(expected).sorted()), str(result)


@pytest.mark.parametrize("case", kat_cases(), ids=lambda c: c["name"])
def test_reference_known_answers(engine, case):
    scs_test(case["trees"], case["expected"], engine, **case["kwargs"])
    scs_test(case["trees"], case["expected"], engine, random_state=np.random.RandomState(7), **case["kwargs"])


@pytest.mark.parametrize("name", ["dcm", "dcm_iq", "supertriplets"])
def test_reference_fixtures(engine, name):
    fixture = json.loads((GOLDEN / f"fixture_{name}.json").read_text())
    result = construct_supertree(parse(fixture["trees"]), pcg_weighting=fixture["weighting"], engine=engine)
    assert rf(result, make_tree(fixture["expected"])) == 0


@pytest.mark.parametrize(
    "name",
    ["dcm", "dcm_iq", "supertriplets", "c1_100x30_depth", "c2_500x50_branch", "c3_1000x100_branch_weighted",
     "s_200x40_bootstrap", "s_300x40_branch_weighted", "s_150x40_one"],
)  # fmt: skip
def test_node_by_node_against_reference_trace(engine, name):
    case = load_case(name)
    trace: list = []
    tree = construct_supertree(
        parse(case["lines"]), case["weights"], case["weighting"], engine=engine, trace=trace
    )
    report = compare_with_reference_trace(trace, case["nodes"], name)
    by_names = {tuple(r["names"]): r for r in case["nodes"]}
    for rec in trace:
        ref = by_names.get(tuple(rec["names"]))
        if ref is None or "eigenvalues" not in ref or rec["contracted_size"] < 3:
            continue
        assert abs(rec["stats"]["eig"][1] - ref["eigenvalues"][1]) < 1e-6, rec["names"]  # Fiedler eigenvalue: 1e-6
    assert sorted(tree.get_tip_names()) == case["names"]
    reference = make_tree(case["supertree"])
    divergent = report.pop("divergent_sets")
    # the supertrees may differ only below a recorded divergence (RF = 0 when there is none)
    assert rf_outside(tree, reference, divergent) == 0
    if not divergent:
        assert len(trace) == len(case["nodes"])
        assert rf(tree, reference) == 0
    print(name, {k: v for k, v in report.items() if k != "divergent_nodes"}, "RF", rf(tree, reference))


def test_argument_errors(engine):
    with pytest.raises(ValueError, match="at least one tree"):
        construct_supertree([], engine=engine)
    with pytest.raises(ValueError, match="Invalid weighting strategy selected: 'bogus'"):
        construct_supertree([make_tree("(a,b)")], pcg_weighting="bogus", engine=engine)
    with pytest.raises(ValueError, match=r"The number of trees \(2\) and tree weights \(1\) must match."):
        construct_supertree([make_tree("(a,b)"), make_tree("(b,c)")], weights=[1.0], engine=engine)
    with pytest.raises(ValueError, match="at least one tree"):
        construct_supertree([NotCompleted()], engine=engine)
    with pytest.raises(TypeError):
        construct_supertree([make_tree("(a,(b,(c,d)))"), make_tree("(a,(c,(b,d)))")], pcg_weighting="bootstrap",
                            engine=engine)  # fmt: skip


def test_not_completed(engine):
    """ref: tests/test_spectral_cluster_supertree.py:258-274"""
    t1, t2 = "(a,(b,(c,d)))", "(a,(c,(b,d)))"
    nc = NotCompleted("ERROR", "test", "failed upstream")
    for trees, weights, expected in [
        ([make_tree(t1), nc, make_tree(t2)], [3, 100, 1], t1),
        ([make_tree(t1), nc, make_tree(t2)], [1, 100, 3], t2),
        ([nc, make_tree(t1)], None, t1),
    ]:
        result = construct_supertree(trees, weights, engine=engine)
        assert result.sorted().same_shape(make_tree(expected).sorted())


def test_load_trees_and_cli(engine, tmp_path, monkeypatch):
    """ref: tests/test_cli.py:14-49"""
    from click.testing import CliRunner

    from spectralclustersupertree_b200 import engine as engine_mod
    from spectralclustersupertree_b200.cli import scs

    monkeypatch.setattr(engine_mod, "_DEFAULT", engine)
    fixture = json.loads((GOLDEN / "fixture_supertriplets.json").read_text())
    in_file = tmp_path / "in.tre"
    out_file = tmp_path / "out.tre"
    in_file.write_text("\n".join(fixture["trees"]) + "\n")
    loaded = load_trees(in_file)
    assert len(loaded) == len(fixture["trees"])
    runner = CliRunner()
    result = runner.invoke(scs, ["-i", str(in_file), "-o", str(out_file), "-p", "DEPTH"])
    assert result.exit_code == 0, result.output
    assert rf(make_tree(out_file.read_text().strip()), make_tree(fixture["expected"])) == 0
    assert runner.invoke(scs, []).exit_code in (0, 2)  # no_args_is_help
    assert "version" in runner.invoke(scs, ["--version"]).output.lower()


@pytest.mark.parametrize("name", ["supertriplets", "s_300x40_branch_weighted", "s_200x40_bootstrap"])
def test_small_node_path_agrees_with_staged_path(engine, name):
    """The one-CTA path (dense Jacobi) and the staged path (union-find, max-merge, Lanczos) must give
    the same components, contraction sizes, eigenvalues and partitions on every recursion node."""
    case = load_case(name)
    fused: list = []
    staged: list = []
    engine.set_small_node_limit(64)  # the one-CTA path at its full capacity (the library's default limit is 32)
    try:
        construct_supertree(parse(case["lines"]), case["weights"], case["weighting"], engine=engine, trace=fused)
        engine.set_small_node_limit(0)
        construct_supertree(parse(case["lines"]), case["weights"], case["weighting"], engine=engine, trace=staged)
    finally:
        engine.set_small_node_limit(32)
    by_names = {tuple(r["names"]): r for r in staged}
    small_nodes = 0
    for rec in fused:
        other = by_names.get(tuple(rec["names"]))
        if other is None:
            continue  # below a tie where the two solvers may legitimately differ
        assert rec["n_components"] == other["n_components"]
        if "partition" not in rec:
            continue
        assert rec["contracted_size"] == other["contracted_size"]
        a, b = rec["stats"], other["stats"]
        if rec["contracted_size"] >= 3:
            small_nodes += a["solver"] == 2
            assert abs(a["eig"][1] - b["eig"][1]) < 1e-9
            assert a["residual"] < 1e-10
        tie = (a["tie_flag"] | b["tie_flag"]) & 3
        if not tie:
            assert {frozenset(p) for p in rec["partition"]} == {frozenset(p) for p in other["partition"]}
            assert a["kmeans_stable_splits"] == b["kmeans_stable_splits"]
    assert small_nodes >= 10


@pytest.mark.parametrize("name", ["dcm", "supertriplets", "c2_500x50_branch", "s_200x40_bootstrap", "s_150x40_one"])
def test_native_driver_matches_python_recursion(engine, name):
    """csrc/driver.cu (breadth-first, batched small nodes) against the per-node Python recursion."""
    from spectralclustersupertree_b200.engine import Forest
    from spectralclustersupertree_b200.scs import supertree_of_forest

    case = load_case(name)
    trees = parse(case["lines"])
    a_trace: list = []
    b_trace: list = []
    a = supertree_of_forest(Forest.from_trees(trees, case["weights"], case["names"]), case["weighting"],
                            engine=engine, trace=a_trace, native=True)  # fmt: skip
    b = supertree_of_forest(Forest.from_trees(trees, case["weights"], case["names"]), case["weighting"],
                            engine=engine, trace=b_trace, native=False)  # fmt: skip
    by_names = {tuple(r["names"]): r for r in b_trace}
    ties = 0
    for rec in a_trace:
        other = by_names.get(tuple(rec["names"]))
        if other is None:
            continue
        assert rec["n_components"] == other["n_components"]
        if "partition" in rec:
            assert rec["contracted_size"] == other["contracted_size"]
            same = {frozenset(p) for p in rec["partition"]} == {frozenset(p) for p in other["partition"]}
            if not same:
                assert (rec["stats"]["tie_flag"] | other["stats"]["tie_flag"]) & 3
                ties += 1
    if ties == 0:
        assert len(a_trace) == len(b_trace)
        assert rf(a, b) == 0
    assert sorted(a.get_tip_names()) == sorted(b.get_tip_names())


def test_batched_small_nodes_are_bit_exact(engine):
    """Graphs built inside small_batch_kernel equal the ones pcg_rows_kernel builds: same eigenvalues to
    rounding, same partitions, on every small node of a job (native driver vs per-node path with the
    staged graph build)."""
    case = load_case("s_300x40_branch_weighted")
    trees = parse(case["lines"])
    fused: list = []
    construct_supertree(trees, case["weights"], case["weighting"], engine=engine, trace=fused)
    ref = load_case("s_300x40_branch_weighted")["nodes"]
    report = compare_with_reference_trace(fused, ref, "s_300x40_branch_weighted")
    assert report["compared"] >= 100


@pytest.mark.parametrize(("name", "world"), [("c2_500x50_branch", 2), ("c2_500x50_branch", 4), ("supertriplets", 3)])
def test_sharded_build_joins_to_the_same_supertree(engine, name, world):
    """scs_supertree_build_sharded: the ranks' outputs (computed here one after the other on one GPU)
    concatenate past the shared prefix into the supertree of the unsharded build."""
    from spectralclustersupertree_b200.engine import Forest, merge_sharded
    from spectralclustersupertree_b200.scs import _tree_from_flat

    case = load_case(name)
    trees = parse(case["lines"])
    forest = Forest.from_trees(trees, case["weights"], case["names"])
    whole = engine.supertree_build(forest, case["weighting"])
    parts = []
    nodes = 0
    for rank in range(world):
        built = engine.supertree_build(forest, case["weighting"], rank=rank, world=world)
        parts.append((built["parent"], built["taxon"], built["shared_prefix"]))
        nodes += built["nodes_small"] + built["nodes_large"]
    assert len({p for _, _, p in parts}) == 1
    parent, taxon = merge_sharded(parts)
    assert len(parent) == len(whole["parent"])
    assert (parent[1:] < np.arange(1, len(parent))).all()
    a = _tree_from_flat(parent, taxon, case["names"])
    b = _tree_from_flat(whole["parent"], whole["taxon"], case["names"])
    assert sorted(a.get_tip_names()) == sorted(b.get_tip_names())
    assert rf(a, b) == 0
    # the replicated waves are solved by every rank, the rest exactly once
    assert nodes >= whole["nodes_small"] + whole["nodes_large"]


@pytest.mark.parametrize("name", ["c2_500x50_branch", "c3_1000x100_branch_weighted", "s_300x40_branch_weighted",
                                  "s_200x40_bootstrap", "dcm"])  # fmt: skip
def test_medium_batch_matches_per_node_path(engine, name):
    """csrc/medium.cu (all nodes of a wave between 65 and 4096 taxa in one batch, Lanczos in lock-step) against the
    per-node staged path: same components, contraction sizes and partitions on every recursion node, Fiedler
    eigenvalues to 1e-9."""
    from spectralclustersupertree_b200.engine import Forest

    case = load_case(name)
    trees = parse(case["lines"])

    def build():
        forest = Forest.from_trees(trees, case["weights"], case["names"])
        return engine.supertree_build(forest, case["weighting"], record=True)

    batched = build()
    engine.set_medium_node_limit(0)
    try:
        staged = build()
    finally:
        engine.set_medium_node_limit(4096)
    assert staged["nodes_medium"] == 0
    by_taxa = {taxa.tobytes(): (part, stats) for taxa, part, stats in staged["records"]}
    compared = 0
    for taxa, part, stats in batched["records"]:
        if len(taxa) <= 64:
            continue
        other = by_taxa.get(taxa.tobytes())
        if other is None:
            continue  # below a tie where the two solvers may legitimately differ
        compared += 1
        opart, ostats = other
        assert stats.n_components == ostats.n_components
        assert stats.contracted_size == ostats.contracted_size
        if stats.n_components != 1:
            assert np.array_equal(part, opart)
            continue
        if stats.contracted_size >= 3:
            assert abs(stats.eig[1] - ostats.eig[1]) < 1e-9
            assert stats.residual < 1e-10
        if not ((stats.tie_flag | ostats.tie_flag) & 3):
            assert np.array_equal(part, opart) or np.array_equal(part, 1 - opart)
            assert stats.kmeans_stable_splits == ostats.kmeans_stable_splits
    if name != "dcm":
        assert batched["nodes_medium"] > 0 and compared > 0
    assert int((batched["taxon"] >= 0).sum()) == int((staged["taxon"] >= 0).sum())


@pytest.mark.parametrize("name", ["c2_500x50_branch", "c3_1000x100_branch_weighted", "s_200x40_bootstrap"])
def test_triangle_graph_build_equals_full_rows(engine, name):
    """The graph build visits every leaf pair once (row CTA a: the columns beyond a; pcg_mirror_bits and pcg_degree_rows: the other
    triangle of W and of the bit matrices, the row sums).  With scs_ctx_set_full_rows every row CTA visits all its
    pairs instead: the whole recursion -- batched medium nodes and per-node path -- must come out identical to the
    last bit of every Fiedler eigenvalue, because W, the bit matrices and the degrees are."""
    from spectralclustersupertree_b200.engine import Forest

    case = load_case(name)
    trees = parse(case["lines"])

    def build():
        forest = Forest.from_trees(trees, case["weights"], case["names"])
        return engine.supertree_build(forest, case["weighting"], record=True)

    outs = []
    for medium_limit in (4096, 0):
        engine.set_medium_node_limit(medium_limit)
        try:
            mirrored = build()
            engine.set_full_rows(True)
            try:
                full = build()
            finally:
                engine.set_full_rows(False)
        finally:
            engine.set_medium_node_limit(4096)
        outs.append(mirrored)
        assert len(mirrored["records"]) == len(full["records"])
        for (taxa, part, stats), (otaxa, opart, ostats) in zip(mirrored["records"], full["records"], strict=True):
            assert np.array_equal(taxa, otaxa)
            assert np.array_equal(part, opart)
            assert stats.n_components == ostats.n_components and stats.contracted_size == ostats.contracted_size
            if stats.n_components == 1 and stats.contracted_size >= 3:
                assert stats.eig[1] == ostats.eig[1] and stats.eig[2] == ostats.eig[2]
                assert stats.matvecs == ostats.matvecs
        assert np.array_equal(mirrored["parent"], full["parent"]) and np.array_equal(mirrored["taxon"], full["taxon"])
    assert outs[0]["nodes_medium"] > 0 and outs[1]["nodes_medium"] == 0


@pytest.mark.parametrize("name", ["dcm", "dcm_iq", "supertriplets", "c2_500x50_branch", "c3_1000x100_branch_weighted",
                                  "s_300x40_branch_weighted", "s_200x40_bootstrap", "s_150x40_one"])  # fmt: skip
def test_device_resident_forest_matches_host_forest(engine, name):
    """csrc/devdriver.cu + devforest.cu (source trees resident in HBM: tours and the restriction to the children of
    every split computed on the device) against csrc/driver.cu + forest.cpp (flat trees restricted by the host
    threads): the same recursion nodes with the same vertex sets, the same partitions, eigenvalues equal to the last
    bits (the restricted branch lengths are the same sums in the same order), the same supertree."""
    from helpers import flat_clades
    from spectralclustersupertree_b200.engine import Forest

    case = load_case(name)
    trees = parse(case["lines"])

    def build():
        forest = Forest.from_trees(trees, case["weights"], case["names"])
        return engine.supertree_build(forest, case["weighting"], record=True)

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
        assert stats.n_components == ostats.n_components
        assert stats.contracted_size == ostats.contracted_size
        assert np.array_equal(part, opart)
        if stats.n_components == 1 and stats.contracted_size >= 3:
            assert stats.eig[1] == ostats.eig[1], (len(taxa), stats.eig[1], ostats.eig[1])
    assert flat_clades(on_device["parent"], on_device["taxon"]) == flat_clades(on_host["parent"], on_host["taxon"])
    assert sorted(on_device["taxon"][on_device["taxon"] >= 0]) == sorted(on_host["taxon"][on_host["taxon"] >= 0])
