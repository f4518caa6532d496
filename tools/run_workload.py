"""One synthetic workload through the product path (scs_forest_create_view + scs_supertree_build over the C ABI),
on 1 GPU or -- under torchrun -- on N GPUs with the large recursion nodes row-sharded over them.

    python tools/run_workload.py c5 [--repeat 2] [--shard-min-n 4096] [--out gpurun_out/c5.json]
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/run_workload.py c5

Prints one JSON line (rank 0): wall time of every repeat, the driver's host/GPU split, per-wave times,
the job's counters, and a checksum of the supertree's clades so that runs at different N can be compared."""

from __future__ import annotations

import os

# idle OpenMP threads of the library's host-side helpers must sleep, not spin: with one process per GPU they would
# take the cores the other ranks' driver threads need (read by the OpenMP runtime when it is first loaded)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import argparse
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200.engine import Engine, Forest, merge_sharded, set_host_threads  # noqa: E402


def clade_checksum(parent: np.ndarray, taxon: np.ndarray) -> tuple[str, int, int]:
    """Order-independent digest of the set of clades (as sets of global taxon ids) of a flat tree."""
    count = len(parent)
    # per-node (sum, xor, size) of 64-bit taxon hashes, accumulated bottom-up (parent < child)
    h = np.zeros(count, dtype=np.uint64)
    x = np.zeros(count, dtype=np.uint64)
    size = np.zeros(count, dtype=np.int64)
    tips = taxon >= 0
    keys = (taxon[tips].astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    keys ^= keys >> np.uint64(29)
    keys *= np.uint64(0xBF58476D1CE4E5B9)
    h[tips] = keys
    x[tips] = keys
    size[tips] = 1
    with np.errstate(over="ignore"):
        for i in range(count - 1, 0, -1):
            p = parent[i]
            h[p] += h[i]
            x[p] ^= x[i]
            size[p] += size[i]
    internal = ~tips
    rows = np.stack([h[internal], x[internal], size[internal].astype(np.uint64)], axis=1)
    rows = rows[np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))]
    return hashlib.sha256(rows.tobytes()).hexdigest()[:16], int(tips.sum()), int(internal.sum())


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("workload", choices=sorted(bench.WORKLOADS))
    parser.add_argument("--repeat", type=int, default=2)
    parser.add_argument("--shard-min-n", type=int, default=4096)
    parser.add_argument("--out", default="")
    args = parser.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("gloo")
        dist = dist_mod
    set_host_threads(max(1, min(16, (os.cpu_count() or 1) // world)))
    t0 = time.perf_counter()
    a = bench.make_workload(args.workload)
    t_make = time.perf_counter() - t0
    engine = Engine(local)
    if dist is not None and args.shard_min_n > 0:
        handles = [None] * world
        dist.all_gather_object(handles, engine.shard_create(rank, world, len(a["names"])))
        engine.shard_connect(handles)
        engine.shard_configure(min_n=args.shard_min_n, timeout_seconds=120.0)
        dist.barrier()
    runs = []
    built = merged = None
    for _ in range(args.repeat):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        forest = Forest.from_arrays(a["node_offsets"], a["parent"], a["length"], a["support"], a["taxon"], a["weights"],
                                    a["names"])  # fmt: skip
        t1 = time.perf_counter()
        built = engine.supertree_build(forest, a["weighting"], rank=rank, world=world)
        t2 = time.perf_counter()
        if dist is not None:
            parts = [None] * world
            dist.all_gather_object(parts, (built["parent"], built["taxon"], built["shared_prefix"]))
            merged = merge_sharded(parts)
        else:
            merged = (built["parent"], built["taxon"])
        runs.append(time.perf_counter() - t0)
        phases = {"forest_create": t1 - t0, "build": t2 - t1, "gather_and_join": time.perf_counter() - t2}
        forest.close()
    digest, tips, internal = clade_checksum(*merged)
    line = {
        "workload": bench.describe(args.workload), "n_gpus": world, "seconds": runs, "phases_last_run": phases,
        "generate_seconds": t_make,
        "host_seconds": built["seconds"], "nodes_small": built["nodes_small"], "nodes_large": built["nodes_large"],
        "waves": built["waves"], "wave_tasks": built["wave_tasks"], "wave_max_n": built["wave_max_n"],
        "wave_seconds_gpu_restrict_total": [[round(1e3 * v, 2) for v in w] for w in built["wave_seconds"]],
        "pair_visits_this_rank": built["pair_visits"], "nodes_row_sharded": engine.shard_nodes if world > 1 else 0,
        "supertree": {"tips": tips, "internal_nodes": internal, "clade_checksum": digest},
    }  # fmt: skip
    if dist is not None:
        engine.synchronize()
        dist.barrier()
    engine.close()
    if rank == 0:
        text = json.dumps(line)
        print(text, flush=True)
        if args.out:
            Path(args.out).write_text(text + "\n")
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
