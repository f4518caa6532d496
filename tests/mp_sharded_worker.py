"""Worker of tests/test_gpu_sharded_nodes.py::test_two_processes_over_cuda_ipc (launched with torchrun,
one process per GPU): row-sharded recursion nodes over CUDA IPC windows against the same nodes on one GPU.

    python -m torch.distributed.run --nproc-per-node 2 tests/mp_sharded_worker.py [n] [trees]
"""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch
    import torch.distributed as dist

    from spectralclustersupertree_b200.engine import Engine, Forest, merge_sharded, set_host_threads
    from spectralclustersupertree_b200.synthetic import make_problem

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    trees = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank)) % max(torch.cuda.device_count(), 1)
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    set_host_threads(max(1, (os.cpu_count() or 1) // world))

    arrays = make_problem(n, trees, "branch", 31337, tree_weights=True).forest_arrays()

    def forest():
        return Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                  arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip

    engine = Engine(local)
    single = engine.supertree_build(forest(), "branch", record=True)

    handle = engine.shard_create(rank, world, n)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    engine.shard_connect(handles)
    engine.shard_configure(min_n=512, timeout_seconds=30.0)
    dist.barrier()

    # one node, cooperatively: the largest connected node of the single-GPU run
    f = forest()
    tours = f.tours("branch")
    part, stats = engine.node_split(tours, seed=9)
    while stats.n_components != 1:
        f = f.induce(f.taxa()[part == np.argmax(np.bincount(part))])
        tours = f.tours("branch")
        part, stats = engine.node_split(tours, seed=9)
    engine.shard_engage(True)
    dist.barrier()
    t0 = time.perf_counter()
    part2, stats2 = engine.node_split(tours, seed=9)
    dt = time.perf_counter() - t0
    engine.shard_engage(False)
    assert engine.shard_nodes == 1
    assert np.array_equal(part, part2), "partition differs from the single-GPU one"
    assert stats.eig[1] == stats2.eig[1] and stats.matvecs == stats2.matvecs, (stats.eig[1], stats2.eig[1])

    # the whole job: large nodes shared out, then the frontier dealt out; the source trees cross PCIe once for all
    # ranks (every rank its slice, the rest gathered from the peers' windows)
    os.environ["SCS_SHARED_UPLOAD_MIN_NODES"] = "0"
    dist.barrier()
    built = engine.supertree_build(forest(), "branch", rank=rank, world=world)
    assert engine.shard_nodes > 1
    parts = [None] * world
    dist.all_gather_object(parts, (built["parent"], built["taxon"], built["shared_prefix"]))
    parent, taxon = merge_sharded(parts)
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import rf
    from spectralclustersupertree_b200.scs import _tree_from_flat

    a = _tree_from_flat(single["parent"], single["taxon"], arrays["names"])
    b = _tree_from_flat(parent, taxon, arrays["names"])
    assert rf(a, b) == 0, "sharded supertree differs from the single-GPU one"
    shared_nodes = engine.shard_nodes
    dist.barrier()
    engine.shard_destroy()
    engine.close()
    if rank == 0:
        print(f"SHARDED-OK world={world} node n={tours.n} m={stats.contracted_size} matvecs={stats.matvecs} "
              f"cooperative node {dt * 1e3:.2f} ms, nodes shared out {shared_nodes}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
