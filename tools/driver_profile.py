"""Breakdown of one native supertree build (run on the GPU box): python tools/driver_profile.py [workload]"""

from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200.engine import Engine, Forest  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    a = bench.make_workload(workload)
    engine = Engine(0)
    for i in range(3):
        t0 = time.perf_counter()
        forest = Forest.from_arrays(a["node_offsets"], a["parent"], a["length"], a["support"], a["taxon"],
                                    a["weights"], a["names"])  # fmt: skip
        t1 = time.perf_counter()
        launches = engine.launch_count
        engine.stage_seconds(reset=True)
        built = engine.supertree_build(forest, a["weighting"])
        stages = engine.stage_seconds(reset=True)
        t2 = time.perf_counter()
        sec = built["seconds"]
        other = (t2 - t1) - sum(sec.values())
        print(f"run {i}: forest_create {t1 - t0:.3f} s, build {t2 - t1:.3f} s = " +
              ", ".join(f"{k} {v:.3f}" for k, v in sec.items()) + f", other {other:.3f}; "
              f"small {built['nodes_small']} large {built['nodes_large']} waves {built['waves']} "
              f"launches {engine.launch_count - launches}")
        print("   staged path: " + ", ".join(f"{k} {v:.3f}" for k, v in stages.items()))
        if i == 2:
            print("   wave: tasks max_n | gpu ms, restrict ms, total ms")
            for w, (tasks, max_n, sec) in enumerate(zip(built["wave_tasks"], built["wave_max_n"], built["wave_seconds"], strict=True)):
                print(f"   {w:3d}: {tasks:5d} {max_n:6d} | {1e3 * sec[0]:7.2f} {1e3 * sec[1]:7.2f} {1e3 * sec[2]:7.2f}")


if __name__ == "__main__":
    main()
