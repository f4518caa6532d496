"""The large recursion nodes of one job, for ncu (run on the GPU box).

    python tools/profile_nodes.py [workload] [min_n] [repeat]

Runs the supertree recursion restricted to nodes with at least ``min_n`` taxa (default 2048): these
are the nodes whose matrices exceed L2, i.e. where the HBM roofline applies.  Prints one line per
node.  Meant to be wrapped by ``ncu`` (launch list, then ``--set full`` on the matvec and the graph
build row kernel); see profiles/README.md."""

from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200.engine import Engine, Forest  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    min_n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    arrays = bench.make_workload(workload)
    engine = Engine(0)
    for _ in range(repeat):
        forest = Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                    arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip
        stack = [forest]
        while stack:
            current = stack.pop()
            if current.num_trees < 2 or len(current.taxa()) < min_n:
                continue
            launches = engine.launch_count
            t0 = time.perf_counter()
            taxa, part, stats = engine.forest_split(current, arrays["weighting"], seed=len(stack))
            dt = time.perf_counter() - t0
            print(f"node n={len(taxa)} T={current.num_trees} L={current.num_leaves} components={stats.n_components} "
                  f"m={stats.contracted_size} matvecs={stats.matvecs} lambda2={stats.eig[1]:.9f} "
                  f"launches={engine.launch_count - launches} {1e3 * dt:.2f} ms", flush=True)
            parts = stats.n_components if stats.n_components != 1 else 2
            for c in range(parts):
                comp = taxa[part == c]
                if len(comp) >= min_n:
                    stack.append(current.induce(comp))
    engine.close()


if __name__ == "__main__":
    main()
