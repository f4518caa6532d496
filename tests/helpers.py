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
    "c3_1000x100_branch_weighted": "branch",
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

# Divergence types.  The first three are the near-ties north_star allows ("partitions identical ... at every
# recursion node whose eigengap exceeds a stated threshold (near-ties reported)"), widened by exactly one case:
# nodes where the REFERENCE ITSELF returns different partitions for different RNG seeds (recorded in the goldens
# as kmeans.seen) and ours is one of its answers.  Everything else is a mismatch and is counted against
# EXPECTED_MISMATCHES.
ALLOWED = ("eigengap_tie", "margin_tie", "kmeans_rng")
MISMATCH = ("kmeans_local_optimum",)

# case -> mismatches known and reported in PARITY.json (node sizes).  s_150x40_one: one node (m = 26,
# lambda_3 - lambda_2 = 0.17) where sklearn's k-means returned the same Lloyd-stable but non-optimal split
# (between-cluster score 0.7 % below the optimum) for all 13 seeds tried; the exact 2-means takes the optimum.
EXPECTED_MISMATCHES = {"s_150x40_one": 1}


def _canon(partition) -> list[str]:
    a, b = sorted(partition[0]), sorted(partition[1])
    return a if a and (not b or a[0] < b[0]) else b


def eig_tie(ref_node: dict) -> bool:
    eig = ref_node.get("eigenvalues")
    return eig is not None and len(eig) >= 3 and eig[2] - eig[1] < GAP_TIE


def classify_divergence(ref_node: dict, ours) -> str | None:
    """Why our bipartition may differ from the reference's at this node: one of ALLOWED / MISMATCH, or None
    if nothing explains it (a bug)."""
    if eig_tie(ref_node):
        return "eigengap_tie"
    margin = ref_node.get("margin")
    if margin is not None and margin < MARGIN_TIE:
        return "margin_tie"
    km = ref_node.get("kmeans") or {}
    mine = _canon(ours)
    if any(mine == seen for seen in km.get("seen", [])):
        return "kmeans_rng"
    if "optimal_partition" in km and mine == _canon(km["optimal_partition"]):
        return "kmeans_local_optimum"
    return None


def is_tie(ref_node: dict) -> bool:
    """Eigengap or margin tie (the two near-tie kinds of the parity contract)."""
    margin = ref_node.get("margin")
    return eig_tie(ref_node) or (margin is not None and margin < MARGIN_TIE)


def compare_with_reference_trace(trace: list[dict], ref_nodes: list[dict], case: str = "") -> dict:
    """Node-by-node comparison on identical vertex sets.

    Every recursion node of our run is looked up in the reference's recorded run by its vertex set and must
    have the same number of components, the same contracted size and the same bipartition (up to label swap).
    A different bipartition must be explained by ``classify_divergence``; nodes below a divergence exist only
    in our run ("orphans") and must lie inside a divergent node's vertex set.  Mismatch-type divergences are
    counted against EXPECTED_MISMATCHES[case].  Returns the counts (also what PARITY.json reports).
    """
    by_names = {tuple(r["names"]): r for r in ref_nodes}
    out = {"compared": 0, "spectral": 0, "orphans": 0, "divergences": {}, "divergent_nodes": []}
    divergent_sets: list[frozenset] = []
    for rec in trace:
        ref = by_names.get(tuple(rec["names"]))
        if ref is None:
            names = frozenset(rec["names"])
            assert any(names <= d for d in divergent_sets), ("node absent from the reference run", rec["names"][:8])
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
        kind = classify_divergence(ref, rec["partition"])
        assert kind is not None, ("partition differs and nothing explains it", rec["names"], ref.get("eigenvalues"))
        out["divergences"][kind] = out["divergences"].get(kind, 0) + 1
        out["divergent_nodes"].append({"n": len(rec["names"]), "m": ref["contracted_size"], "kind": kind,
                                       "eigenvalues": ref.get("eigenvalues"), "margin": ref.get("margin")})  # fmt: skip
        divergent_sets.append(frozenset(rec["names"]))
    out["tie_divergences"] = sum(out["divergences"].get(k, 0) for k in ALLOWED)
    out["mismatches"] = sum(out["divergences"].get(k, 0) for k in MISMATCH)
    assert out["mismatches"] <= EXPECTED_MISMATCHES.get(case, 0), out["divergences"]
    out["divergent_sets"] = divergent_sets
    return out


def rf_outside(tree, reference, divergent_sets) -> int:
    """Robinson-Foulds distance counting only clades that are NOT inside a divergent node's vertex set: must be
    0 -- the two supertrees may only differ below the nodes where a divergence was recorded."""
    diff = tree.clade_sets() ^ reference.clade_sets()
    return sum(1 for clade in diff if not any(clade <= d for d in divergent_sets))


# ---- compact traces of the CPU oracle's whole recursion at full workload size (tools/oracle_run.py) --------
def ctrace_key(ids: np.ndarray) -> str:
    import hashlib

    return hashlib.blake2b(np.ascontiguousarray(ids, dtype="<i4").tobytes(), digest_size=8).hexdigest()


def ctrace_part_hash(labels: np.ndarray) -> str:
    """Hash of a labelling up to renaming of the labels (numbered by first appearance)."""
    import hashlib

    values, first, inverse = np.unique(np.asarray(labels), return_index=True, return_inverse=True)
    rank = np.empty(len(values), dtype=np.int32)
    rank[np.argsort(first)] = np.arange(len(values), dtype=np.int32)
    return hashlib.blake2b(rank[inverse].astype("<i4").tobytes(), digest_size=8).hexdigest()


def load_ctrace(workload: str) -> dict:
    import gzip

    with gzip.open(GOLDEN / f"ctrace_{workload}.json.gz", "rt") as fh:
        return json.load(fh)


def compare_with_ctrace(records, ctrace: dict, expected_mismatches: int = 0) -> dict:
    """Every recursion node of a GPU build (``Engine.supertree_build(record=True)`` records: taxa, part, stats)
    against the oracle's compact trace: same components (count and labelling), same contracted size, Fiedler
    eigenvalue within 1e-6, same bipartition -- or a divergence explained by the same rules as
    ``classify_divergence``.  Returns the counts PARITY.json reports."""
    by_key = {node["key"]: node for node in ctrace["nodes"]}
    out = {"recursion_nodes": len(records), "compared": 0, "spectral": 0, "orphans": 0, "divergences": {},
           "divergent_nodes": [], "max_eig_error": 0.0}  # fmt: skip
    divergent_sets: list[frozenset] = []
    for taxa, part, stats in records:
        ref = by_key.get(ctrace_key(taxa))
        if ref is None:
            mine = frozenset(taxa.tolist())
            assert any(mine <= d for d in divergent_sets), ("node absent from the oracle's run", len(taxa))
            out["orphans"] += 1
            continue
        out["compared"] += 1
        assert int(stats.n_components) == ref["nc"], (len(taxa), stats.n_components, ref["nc"])
        if "m" not in ref:
            assert ctrace_part_hash(part) == ref["part"] if "part" in ref else True
            continue
        out["spectral"] += 1
        assert int(stats.contracted_size) == ref["m"], (len(taxa), stats.contracted_size, ref["m"])
        if "eig" in ref:
            err = abs(stats.eig[1] - ref["eig"][0])
            out["max_eig_error"] = max(out["max_eig_error"], err)
            assert err < 1e-6, (len(taxa), stats.eig[1], ref["eig"][0])  # Fiedler eigenvalue: 1e-6 (north_star)
        mine = ctrace_part_hash(part)
        if mine == ref["part"]:
            continue
        km = ref.get("km") or {}
        if "eig" in ref and ref["eig"][1] - ref["eig"][0] < GAP_TIE:
            kind = "eigengap_tie"
        elif ref.get("margin", 1.0) < MARGIN_TIE:
            kind = "margin_tie"
        elif mine in km.get("seen", []) or mine == ref.get("nat"):
            kind = "kmeans_rng"
        elif mine == km.get("opt"):
            kind = "kmeans_local_optimum"
        else:
            raise AssertionError(("partition differs and nothing explains it", len(taxa), ref))
        out["divergences"][kind] = out["divergences"].get(kind, 0) + 1
        out["divergent_nodes"].append({"n": len(taxa), "m": ref["m"], "kind": kind, "eig": ref.get("eig"),
                                       "margin": ref.get("margin")})  # fmt: skip
        divergent_sets.append(frozenset(taxa.tolist()))
    out["tie_divergences"] = sum(out["divergences"].get(k, 0) for k in ALLOWED)
    out["mismatches"] = sum(out["divergences"].get(k, 0) for k in MISMATCH)
    assert out["mismatches"] <= expected_mismatches, out["divergences"]
    out["divergent_sets"] = divergent_sets
    return out


def flat_clades(parent: np.ndarray, taxon: np.ndarray) -> set[frozenset]:
    """Clade sets (as frozensets of global taxon ids, tips and the root excluded) of a flat supertree."""
    count = len(parent)
    below: list[list[int] | None] = [None] * count
    clades: set[frozenset] = set()
    for i in range(count - 1, -1, -1):
        if taxon[i] >= 0:
            mine = [int(taxon[i])]
        else:
            mine = below[i] or []
            if i > 0 and len(mine) > 1:
                clades.add(frozenset(mine))
        p = parent[i]
        if p >= 0:
            if below[p] is None:
                below[p] = []
            below[p].extend(mine)
        below[i] = None
    return clades
