"""Loaders for the golden vectors under tests/golden (recorded from the unmodified reference by
tests/golden/make_golden.py) and small comparison utilities."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from spectralclustersupertree_b200.tree import make_tree

GOLDEN = Path(__file__).resolve().parent / "golden"

# case name -> weighting (cases with pcg_<case>.npz and trace_<case>.json)
CASES = {
    "dcm": "one",
    "dcm_iq": "branch",
    "supertriplets": "depth",
    "c1_100x30_depth": "depth",
    "c2_500x50_branch": "branch",
    "s_200x40_bootstrap": "bootstrap",
    "s_300x40_branch_weighted": "branch",
    "s_150x40_one": "one",
}


def kat_cases() -> list[dict]:
    return json.loads((GOLDEN / "kat_cases.json").read_text())


def load_case(name: str) -> dict:
    """Trees (Newick lines), weights, weighting, names, reference trace and reference PCG arrays."""
    trace = json.loads((GOLDEN / f"trace_{name}.json").read_text())
    lines = trace["trees"]
    expected = None
    if lines is None:
        fixture = json.loads((GOLDEN / trace["fixture"]).read_text())
        lines = fixture["trees"]
        expected = fixture["expected"]
    weights = trace["weights"]
    if weights is None:
        weights = [1.0] * len(lines)
    pcg = dict(np.load(GOLDEN / f"pcg_{name}.npz"))
    return {
        "name": name,
        "weighting": trace["weighting"],
        "lines": lines,
        "weights": weights,
        "names": trace["names"],
        "nodes": trace["nodes"],
        "supertree": trace["supertree"],
        "expected": expected,
        "pcg": pcg,
    }


def parse(lines):
    return [make_tree(s) for s in lines]


def same_partition(a: list[list[str]], b: list[list[str]]) -> bool:
    """Two bipartitions equal up to label swap."""
    sa = {frozenset(p) for p in a}
    sb = {frozenset(p) for p in b}
    return sa == sb


def rf(a, b) -> int:
    return len(a.clade_sets() ^ b.clade_sets())


GAP_TIE = 1e-7  # lambda_3 - lambda_2 below this: the Fiedler vector is not unique (any split is valid)
MARGIN_TIE = 1e-9  # a vertex this close to the 2-means boundary may fall on either side


def is_tie(ref_node: dict) -> bool:
    """A reference recursion node whose bipartition is not determined by the mathematics:
    a repeated Fiedler eigenvalue, a vertex on the 2-means boundary, or a k-means with several
    Lloyd-stable splits where the reference's own answer depends on its RNG seed or is not the
    optimum (recorded per node by tests/golden/make_golden.py)."""
    eig = ref_node.get("eigenvalues")
    if eig is not None and len(eig) >= 3 and eig[2] - eig[1] < GAP_TIE:
        return True
    margin = ref_node.get("margin")
    if margin is not None and margin < MARGIN_TIE:
        return True
    km = ref_node.get("kmeans")
    return km is not None and km["stable_splits"] > 1 and not (km["reference_is_optimal"] and km["seed_stable"])


def eig_tie(ref_node: dict) -> bool:
    eig = ref_node.get("eigenvalues")
    return eig is not None and len(eig) >= 3 and eig[2] - eig[1] < GAP_TIE


def compare_with_reference_trace(trace: list[dict], ref_nodes: list[dict]) -> dict:
    """Node-by-node comparison on identical vertex sets.

    Returns counts: ``compared`` nodes found in the reference trace, ``spectral`` of them that went
    through the spectral step, ``tie_divergences`` (partition differs at a tie node: allowed,
    reported) and ``orphans`` (nodes below a tie divergence, absent from the reference trace).
    Raises AssertionError on any difference that is not explained by a tie.
    """
    by_names = {tuple(r["names"]): r for r in ref_nodes}
    out = {"compared": 0, "spectral": 0, "tie_divergences": 0, "orphans": 0}
    for rec in trace:
        ref = by_names.get(tuple(rec["names"]))
        if ref is None:
            out["orphans"] += 1
            continue
        out["compared"] += 1
        assert rec["n_components"] == ref["n_components"], rec["names"]
        if "partition" not in ref:
            continue
        out["spectral"] += 1
        assert rec["contracted_size"] == ref["contracted_size"], rec["names"]
        if same_partition(rec["partition"], ref["partition"]):
            continue
        assert is_tie(ref), ("partition differs at a node that is not a tie", rec["names"], ref.get("eigenvalues"))
        km = ref.get("kmeans") or {}
        if "optimal_partition" in km and not (eig_tie(ref)):
            # the reference stopped in a worse local optimum: ours must be the global one
            assert same_partition(rec["partition"], km["optimal_partition"]), rec["names"]
        out["tie_divergences"] += 1
    assert out["orphans"] == 0 or out["tie_divergences"] > 0
    return out
