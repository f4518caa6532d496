"""GPU: whole jobs at BASELINE.json's sizes against the CPU oracle.

* C3 (1 000 taxa x 100 trees, branch weighting, tree weights) and C4 (10 000 x 1 000, depth): every recursion
  node of the native build against the compact trace of the oracle's whole recursion
  (tests/golden/ctrace_<workload>.json.gz, written by tools/oracle_run.py in the build container: 13 s and
  12 min of CPU): components, contracted size, Fiedler eigenvalue (1e-6), bipartition, and the final supertree
  (RF = 0 outside recorded divergences).  C3 is additionally pinned to the unmodified reference's own run
  (tests/test_gpu_supertree.py, trace_c3_1000x100_branch_weighted.json).
* C5 (50 000 x 5 000, branch): rows of the top-level W against the C oracle restricted to those rows,
  bit-exact, and the recursion node's Fiedler eigenvalue against ARPACK iterating on the same operator.
"""

from __future__ import annotations

import numpy as np
import pytest

from helpers import compare_with_ctrace, flat_clades, load_ctrace
from spectralclustersupertree_b200.tree import make_tree

pytestmark = pytest.mark.gpu


def _workload_forest(workload: str):
    import bench
    from spectralclustersupertree_b200.engine import Forest

    arrays = bench.make_workload(workload)
    forest = Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip
    return arrays, forest


@pytest.mark.parametrize("workload", ["c3", "c4"])
def test_every_recursion_node_against_the_oracle_trace(engine, workload):
    ctrace = load_ctrace(workload)
    arrays, forest = _workload_forest(workload)
    assert len(arrays["names"]) == ctrace["names"]
    built = engine.supertree_build(forest, arrays["weighting"], record=True)
    report = compare_with_ctrace(built["records"], ctrace)
    divergent = report.pop("divergent_sets")
    assert report["compared"] + report["orphans"] == len(built["records"])
    if not divergent:
        assert len(built["records"]) == len(ctrace["nodes"])
    # final supertree: clades may differ only below a recorded divergence (RF = 0 when there is none)
    names = arrays["names"]
    gid = {name: i for i, name in enumerate(names)}
    reference = {frozenset(gid[x] for x in clade) for clade in make_tree(ctrace["supertree"]).clade_sets()
                 if 1 < len(clade) < len(names)}  # fmt: skip
    ours = {c for c in flat_clades(built["parent"], built["taxon"]) if len(c) < len(names)}
    outside = [c for c in ours ^ reference if not any(c <= d for d in divergent)]
    assert not outside, (len(outside), len(ours ^ reference))
    assert int((built["taxon"] >= 0).sum()) == len(names)
    print(workload, {k: v for k, v in report.items() if k != "divergent_nodes"}, "RF", len(ours ^ reference))


def test_c5_top_level_rows_and_fiedler_value(engine):
    """C5 (50 000 taxa x 5 000 trees, branch): the 20 GB W of the top-level node cannot be compared whole, so
    (1) blocks of its rows are compared bit for bit with the C oracle restricted to those rows
    (oracle/pcg_oracle.c:pcg_oracle_rows, itself pinned to the dense oracle by tests/test_oracle_golden.py), and
    (2) the Fiedler eigenvalue of the first connected recursion node below it is compared (1e-6) with ARPACK
    (scipy eigsh) iterating on the same device-resident operator through scs_normalized_matvec_dev."""
    import bench
    from oracle import scs_oracle
    from scipy.sparse.linalg import LinearOperator, eigsh
    from spectralclustersupertree_b200.engine import unpack_bits

    arrays, forest = _workload_forest("c5")
    n = len(arrays["names"])
    taxa, part, stats = engine.forest_split(forest, "branch", contract_edges=True, seed=1)
    assert len(taxa) == n
    csr = bench.oracle_children_csr(arrays)
    rng = np.random.RandomState(11)
    for lo in (0, int(rng.randint(1000, n - 1000)), n - 8):
        Wg, bits = engine.last_node_rows(lo, lo + 8)
        Wo, Co = scs_oracle.pcg_rows_c_arrays(n, *csr, arrays["weights"], "branch", lo, lo + 8)
        assert np.array_equal(Wg, Wo), lo  # bit-exact, branch weighting
        assert np.array_equal(unpack_bits(bits, n), Co > 0), lo
    # walk down the largest component until a node goes through the spectral step
    sub = forest
    for _ in range(12):
        if stats.n_components == 1:
            break
        sub = sub.induce(taxa[part == np.argmax(np.bincount(part))])
        taxa, part, stats = engine.forest_split(sub, "branch", contract_edges=True, seed=1)
    assert stats.n_components == 1 and stats.solver == 3
    p = engine.last_node_pointers()
    m = p["m"] if p["m"] else p["n"]
    W_dev = p["Wc"] if p["m"] and p["m"] != p["n"] else p["W"]
    ones = np.ones(m)
    d_ones = engine.to_device(ones)
    try:
        degree = engine.matvec_on_device(m, W_dev, d_ones, ones)  # isd = 1, x = 1: the row sums
    finally:
        engine.free(d_ones)
    isd = np.where(degree > 0, 1.0 / np.sqrt(np.where(degree > 0, degree, 1.0)), 1.0)
    d_isd = engine.to_device(isd)
    try:
        op = LinearOperator((m, m), matvec=lambda x: engine.matvec_on_device(m, W_dev, d_isd, x), dtype=np.float64)
        vals = eigsh(op, k=3, which="LA", tol=1e-12, v0=np.ones(m), return_eigenvectors=False)
    finally:
        engine.free(d_isd)
    vals = np.sort(vals)[::-1]  # 1 (trivial), then 1 - lambda_2
    assert abs(vals[0] - 1.0) < 1e-9
    assert abs((1.0 - vals[1]) - stats.eig[1]) < 1e-6, (1.0 - vals[1], stats.eig[1])
    print("c5: n", n, "spectral node m", m, "lambda2", stats.eig[1], "ARPACK", 1.0 - vals[1], "matvecs", stats.matvecs)
