"""Host-side costs of the recursion driver on the box it runs on: forest restriction and tour flattening
by thread count (python tools/host_profile.py [workload])."""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200.engine import Forest, set_host_threads  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    a = bench.make_workload(workload)
    f = Forest.from_arrays(a["node_offsets"], a["parent"], a["length"], a["support"], a["taxon"], a["weights"], a["names"])
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "nodes", f.num_nodes, "leaves",
          f.num_leaves, "trees", f.num_trees)
    rng = np.random.RandomState(0)
    keep = np.flatnonzero(rng.rand(len(a["names"])) < 0.87).astype(np.int32)
    for threads in (1, 2, 4, 8, 16, 32, 64):
        if threads > (os.cpu_count() or 1):
            break
        set_host_threads(threads)
        induce, tours = [], []
        for _ in range(8):
            t0 = time.perf_counter()
            g = f.induce(keep)
            induce.append(time.perf_counter() - t0)
            g.close()
            t0 = time.perf_counter()
            f.tours(a["weighting"])
            tours.append(time.perf_counter() - t0)
        print(f"threads {threads:3d}: induce ms " + " ".join(f"{1e3 * x:6.2f}" for x in induce) +
              " | tours ms " + " ".join(f"{1e3 * x:6.2f}" for x in tours))


if __name__ == "__main__":
    main()
