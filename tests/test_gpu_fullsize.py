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
