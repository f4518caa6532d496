"""Wall time of the Python drop-in ``construct_supertree(list[PhyloNode])`` on a bench workload (default c4), with the
host-side share (flattening the node objects) printed beside it.  usage: python tools/python_api_time.py [c4]"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import bench
    from spectralclustersupertree_b200 import construct_supertree
    from spectralclustersupertree_b200.engine import Engine, Forest
    from spectralclustersupertree_b200.synthetic import make_problem

    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    n, t, weighting, seed, tw = bench.WORKLOADS[workload]
    problem = make_problem(n, t, weighting, seed, tree_weights=tw)
    objects = problem.phylonodes()
    weights = problem.weights
    with Engine(0) as engine:
        construct_supertree(objects, weights, weighting, engine=engine)  # warm
        times, flatten = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            result = construct_supertree(objects, weights, weighting, engine=engine)
            times.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            Forest.from_trees(objects, [1.0] * len(objects)).close()
            flatten.append(time.perf_counter() - t0)
    print(json.dumps({"workload": bench.describe(workload), "construct_supertree_s": times, "of_which_flatten_s": flatten,
                      "tips": len(result.get_tip_names())}))  # fmt: skip


if __name__ == "__main__":
    main()
