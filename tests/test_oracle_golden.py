"""CPU: the oracle (oracle/scs_oracle.py + oracle/pcg_oracle.c) against the golden vectors recorded
from the reference's own functions, and against the reference's known-answer cases and fixtures
(ref: tests/test_spectral_cluster_supertree.py, tests/test_data/*.tre)."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import CASES, compare_with_reference_trace, kat_cases, load_case, parse, rf, rf_outside
from oracle import scs_oracle
from spectralclustersupertree_b200.tree import make_tree

SMALL = ["supertriplets", "c1_100x30_depth", "s_150x40_one", "s_200x40_bootstrap"]


@pytest.mark.parametrize("name", list(CASES))
def test_pcg_c_oracle_matches_reference(name):
    case = load_case(name)
    trees = parse(case["lines"])
    tid = {x: i for i, x in enumerate(case["names"])}
    W, C, occ = scs_oracle.pcg_dense_c(trees, case["weights"], case["weighting"], tid)
    ref = case["pcg"]
    assert np.array_equal(W, ref["W"])  # bit-exact, every weighting: same summation order
    assert np.array_equal(C, ref["C"])
    assert np.array_equal(occ, ref["occ"])


@pytest.mark.parametrize("name", SMALL)
def test_pcg_python_oracle_matches_reference(name):
    case = load_case(name)
    trees = parse(case["lines"])
    tid = {x: i for i, x in enumerate(case["names"])}
    W, C, occ = scs_oracle.pcg_dense(trees, case["weights"], case["weighting"], tid)
    ref = case["pcg"]
    assert np.array_equal(W, ref["W"])
    assert np.array_equal(C, ref["C"])
    assert np.array_equal(occ, ref["occ"])


@pytest.mark.parametrize("name", list(CASES))
def test_components_and_contraction_match_reference(name):
    ref = load_case(name)["pcg"]
    label = scs_oracle.graph_components(ref["C"] > 0)
    assert np.array_equal(label, ref["label"])
    group, Wc, Ac = scs_oracle.contract_dense(ref["W"], ref["C"], ref["occ"])
    assert np.array_equal(group, ref["group"])
    m = int(group.max()) + 1
    assert np.array_equal(Wc, ref["Wc"][:m, :m])
    assert np.array_equal(Ac, ref["Ac"][:m, :m])


@pytest.mark.parametrize("case", kat_cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_kats(case):
    for seed in range(2):
        result = scs_oracle.construct_supertree(
            parse(case["trees"]), random_state=np.random.RandomState(seed), **case["kwargs"]
        )
        assert result.sorted().same_shape(make_tree(case["expected"]).sorted()), str(result)


@pytest.mark.parametrize("name", ["dcm", "dcm_iq", "supertriplets", "c1_100x30_depth"])
def test_oracle_recursion_matches_reference_trace(name):
    case = load_case(name)
    trace: list = []
    tree = scs_oracle.construct_supertree(
        parse(case["lines"]), case["weights"], case["weighting"], random_state=np.random.RandomState(0),
        use_c=True, trace=trace,
    )  # fmt: skip
    report = compare_with_reference_trace(trace, case["nodes"], name)
    assert rf_outside(tree, make_tree(case["supertree"]), report["divergent_sets"]) == 0
    if not report["divergent_sets"]:
        assert len(trace) == len(case["nodes"])
        assert rf(tree, make_tree(case["supertree"])) == 0
    if case["expected"] is not None:
        assert rf(tree, make_tree(case["expected"])) == 0


def test_oracle_eigenvalues_match_trace():
    case = load_case("supertriplets")
    ref = case["pcg"]
    # top-level graph of supertriplets is disconnected; use the first spectral node of the trace instead
    node = next(r for r in case["nodes"] if "eigenvalues" in r)
    names = node["names"]
    trees = parse(case["lines"])
    sub = [t.get_sub_tree(names, ignore_missing=True, as_rooted=True) for t in trees
           if len(set(names) & set(t.get_tip_names())) >= 2]  # fmt: skip
    tid = {x: i for i, x in enumerate(sorted(names))}
    W, C, occ = scs_oracle.pcg_dense_c(sub, [1.0] * len(sub), case["weighting"], tid)
    _, Wc, _ = scs_oracle.contract_dense(W, C, occ)
    vals, _ = scs_oracle.normalized_affinity_eigs(Wc, 3)
    assert np.allclose(vals, node["eigenvalues"], atol=1e-10)
    assert ref["W"].shape[0] == len(case["names"])


@pytest.mark.parametrize("name", ["c2_500x50_branch", "s_200x40_bootstrap", "supertriplets"])
def test_row_block_oracle_equals_the_dense_oracle(name):
    """oracle/pcg_oracle.c:pcg_oracle_rows (tip-wise, a block of rows) against pcg_oracle_dense (node-wise, the whole
    matrix) and against the reference's own W: bit for bit."""
    case = load_case(name)
    trees = parse(case["lines"])
    tid = {x: i for i, x in enumerate(case["names"])}
    arrs = scs_oracle.children_csr(trees, tid, case["weighting"])
    n = len(tid)
    for lo, hi in [(0, 7), (n // 2, n // 2 + 13), (n - 5, n)]:
        Wr, Cr = scs_oracle.pcg_rows_c_arrays(n, *arrs, case["weights"], case["weighting"], lo, hi)
        assert np.array_equal(Wr, case["pcg"]["W"][lo:hi])
        assert np.array_equal(Cr, case["pcg"]["C"][lo:hi])


@pytest.mark.parametrize("case", ["bootstrap", "depth", "nocontract"])
def test_untidy_traces_are_reproduced_by_the_oracle(case):
    """tests/golden/ctrace_untidy_*.json.gz (source trees with unary chains, polytomies, missing lengths, unary
    roots; made by tests/golden/make_untidy.py) pin what the GPU tests compare with: the oracle run here must give
    the same recursion nodes, components, contracted sizes and partitions again."""
    import sys

    from pathlib import Path

    from helpers import load_ctrace

    tools = str(Path(__file__).resolve().parent.parent / "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)
    from oracle_run import trace_recursion

    ctrace = load_ctrace(f"untidy_{case}")
    trees = parse(ctrace["lines"])
    names = sorted({x for t in trees for x in t.get_tip_names()})
    assert ctrace["unary_nodes"] > 50 and ctrace["polytomies"] > 10
    _, records, tree = trace_recursion(trees, ctrace["weights"], ctrace["weighting"], names, steer=True,
                                       seeds=ctrace["seeds"], contract_edges=ctrace.get("contract_edges", True))  # fmt: skip
    assert len(records) == len(ctrace["nodes"])
    for got, want in zip(records, ctrace["nodes"], strict=True):
        assert (got["key"], got["n"], got["nc"], got.get("m"), got.get("part")) == (
            want["key"], want["n"], want["nc"], want.get("m"), want.get("part"))  # fmt: skip
        if "eig" in want:
            assert abs(got["eig"][0] - want["eig"][0]) < 1e-9
    assert tree.clade_sets() == make_tree(ctrace["supertree"]).clade_sets()
