"""Where the time of one supertree job goes, by recursion-node size (run on the GPU box).

    python tools/node_profile.py [workload]

Wall-clocks every ``forest_split`` (one recursion node through the C ABI, host buffers in/out) and
the host-side restriction (``Forest.induce``), then prints totals per size bucket."""

from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200 import engine as engine_mod  # noqa: E402
from spectralclustersupertree_b200.engine import Engine, Forest  # noqa: E402
from spectralclustersupertree_b200.scs import supertree_of_forest  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    arrays = bench.make_workload(workload)
    engine = Engine(0)
    forest = Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip
    supertree_of_forest(forest, arrays["weighting"], engine=engine)  # warm-up (allocations)

    records = []
    induce_s = [0.0]
    real_split = Engine.forest_split
    real_induce = Forest.induce

    def timed_split(self, f, weighting, contract_edges=True, seed=0):
        launches = self.launch_count
        t0 = time.perf_counter()
        out = real_split(self, f, weighting, contract_edges=contract_edges, seed=seed)
        dt = time.perf_counter() - t0
        stats = out[2]
        records.append((len(out[0]), f.num_trees, f.num_leaves, stats.n_components, stats.contracted_size,
                        stats.matvecs, dt, self.launch_count - launches))  # fmt: skip
        return out

    def timed_induce(self, keep):
        t0 = time.perf_counter()
        out = real_induce(self, keep)
        induce_s[0] += time.perf_counter() - t0
        return out

    engine_mod.Engine.forest_split = timed_split
    engine_mod.Forest.induce = timed_induce
    forest = Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip
    t0 = time.perf_counter()
    supertree_of_forest(forest, arrays["weighting"], engine=engine)
    total = time.perf_counter() - t0
    rec = np.array(records, dtype=np.float64)
    print(f"workload {workload}: total {total:.3f} s, nodes {len(rec)}, in forest_split {rec[:, 6].sum():.3f} s, "
          f"in induce {induce_s[0]:.3f} s, other host {total - rec[:, 6].sum() - induce_s[0]:.3f} s")
    edges = [0, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 1 << 30]
    print(f"{'n range':>14} {'nodes':>6} {'spectral':>8} {'sum ms':>9} {'ms/node':>8} {'matvecs':>8} {'launches':>9} {'leaves':>10}")
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (rec[:, 0] >= lo) & (rec[:, 0] < hi)
        if not sel.any():
            continue
        r = rec[sel]
        print(f"{lo:>6}-{hi if hi < 1 << 30 else 'inf':>7} {len(r):>6} {int((r[:, 3] == 1).sum()):>8} {1e3 * r[:, 6].sum():>9.1f} "
              f"{1e3 * r[:, 6].mean():>8.3f} {int(r[:, 5].sum()):>8} {int(r[:, 7].sum()):>9} {int(r[:, 2].sum()):>10}")
    big = rec[np.argsort(-rec[:, 6])[:8]]
    print("slowest nodes: n, T, L, ncomp, m, matvecs, ms, launches")
    for r in big:
        print("  ", int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), f"{1e3 * r[6]:.2f}", int(r[7]))


if __name__ == "__main__":
    main()
