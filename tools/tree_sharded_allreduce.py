"""Tree-sharded W accumulation + NCCL all-reduce beside the row-sharded build (BASELINE.json config 4).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/tree_sharded_allreduce.py [c4]

Top-level recursion node of the workload, graph-build stage only, on N GPUs:
  A  tree-sharded: rank r builds a full n x n partial W from its contiguous block of the source trees
     (scs_pcg_build_dev), then W is summed over the ranks with an NCCL all-reduce (and the adjacency bits, occ and
     degree would follow the same way: timed for W, the dominant 8 n^2 bytes);
  B  row-sharded (what the engine does): every rank builds rows [r n/N, (r+1) n/N) of W from ALL trees
     (scs_pcg_build_rows_dev); nothing is exchanged for W, every entry keeps its tree-ordered sum.
Both are timed with CUDA events (max over ranks), 3 warm-ups + 10 timed repetitions; A is also checked against B
(equal for integer-valued weights; for branch weights the order of the additions differs).
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch
    import torch.distributed as dist

    import bench
    from spectralclustersupertree_b200 import _lib
    from spectralclustersupertree_b200.engine import Engine, Forest, set_host_threads

    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    set_host_threads(max(1, (os.cpu_count() or 1) // world))
    arrays = bench.make_workload(workload)
    forest = Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip
    tours = forest.tours(arrays["weighting"])
    n, T = tours.n, tours.num_trees
    lib = _lib.load()
    engine = Engine(local)
    device = torch.device("cuda", local)
    words = lib.scs_bit_words(n)

    def upload(array):
        return torch.from_numpy(np.ascontiguousarray(array)).to(device)

    # A: this rank's block of trees as a tour of its own
    t0, t1 = rank * T // world, (rank + 1) * T // world
    l0, l1 = int(tours.leaf_offsets[t0]), int(tours.leaf_offsets[t1])
    part = {
        "off": upload((tours.leaf_offsets[t0 : t1 + 1] - l0).astype(np.int64)), "tax": upload(tours.leaf_taxon[l0:l1]),
        "dep": upload(tours.adj_depth[l0:l1]), "val": upload(tours.adj_val[l0:l1]),
        "root": upload(tours.root_depth[t0:t1]), "w": upload(tours.tree_weight[t0:t1]),
    }  # fmt: skip
    full = {"off": upload(tours.leaf_offsets.astype(np.int64)), "tax": upload(tours.leaf_taxon), "dep": upload(tours.adj_depth),
            "val": upload(tours.adj_val), "root": upload(tours.root_depth), "w": upload(tours.tree_weight)}  # fmt: skip
    W_part = torch.empty((n, n), dtype=torch.float64, device=device)
    occ = torch.empty(n, dtype=torch.int32, device=device)
    bits = torch.empty((n, words), dtype=torch.int32, device=device)
    mbits = torch.empty((n, words), dtype=torch.int32, device=device)
    degree = torch.empty(n, dtype=torch.float64, device=device)
    r0, r1 = rank * n // world, (rank + 1) * n // world
    W_rows = torch.empty((r1 - r0, n), dtype=torch.float64, device=device)

    def build_tree_block():
        d = part
        status = lib.scs_pcg_build_dev(engine.handle, n, t1 - t0, l1 - l0, d["off"].data_ptr(), d["tax"].data_ptr(),
                                       d["dep"].data_ptr(), d["val"].data_ptr(), d["root"].data_ptr(), d["w"].data_ptr(),
                                       W_part.data_ptr(), None, occ.data_ptr(), bits.data_ptr(), mbits.data_ptr(),
                                       degree.data_ptr())  # fmt: skip
        assert status == 0, status

    def build_row_block():
        d = full
        status = lib.scs_pcg_build_rows_dev(engine.handle, n, T, int(tours.leaf_offsets[-1]), d["off"].data_ptr(),
                                            d["tax"].data_ptr(), d["dep"].data_ptr(), d["val"].data_ptr(),
                                            d["root"].data_ptr(), d["w"].data_ptr(), r0, r1, W_rows.data_ptr(),
                                            occ.data_ptr(), bits.data_ptr(), mbits.data_ptr(), degree.data_ptr())  # fmt: skip
        assert status == 0, status

    def timed(fn, reps=10, warm=3):
        times = []
        for i in range(warm + reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            engine.timer_start()
            fn()
            ms = engine.timer_stop()
            if i >= warm:
                times.append(ms)
        return float(np.mean(times))

    def timed_allreduce(reps=10, warm=3):
        times = []
        for i in range(warm + reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            if world > 1:
                dist.all_reduce(W_part)
            stop.record()
            stop.synchronize()
            if i >= warm:
                times.append(start.elapsed_time(stop))
        return float(np.mean(times))

    build_a = timed(build_tree_block)
    build_b = timed(build_row_block)
    build_tree_block()
    engine.synchronize()
    reduce_a = timed_allreduce()
    # correctness: a fresh partial + one all-reduce against the row block
    build_tree_block()
    build_row_block()
    engine.synchronize()
    if world > 1:
        dist.all_reduce(W_part)
    torch.cuda.synchronize()
    same = bool(torch.equal(W_part[r0:r1], W_rows))
    close = bool(torch.allclose(W_part[r0:r1], W_rows, rtol=1e-12, atol=0))
    numbers = torch.tensor([build_a, reduce_a, build_b], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(numbers, op=dist.ReduceOp.MAX)
    build_a, reduce_a, build_b = numbers.tolist()
    if rank == 0:
        print(json.dumps({
            "workload": bench.describe(workload), "n_gpus": world, "n": n, "trees": T,
            "tree_sharded": {"build_partial_W_ms": build_a, "nccl_allreduce_W_ms": reduce_a, "total_ms": build_a + reduce_a,
                             "allreduce_bytes": 8 * n * n, "busbw_GBps": (2 * (world - 1) / world * 8 * n * n / (reduce_a * 1e-3) / 1e9) if world > 1 and reduce_a > 0 else None},
            "row_sharded": {"build_row_block_ms": build_b, "exchanged_bytes_for_W": 0},
            "W_equal_bit_for_bit": same, "W_equal_to_1e-12": close,
        }), flush=True)  # fmt: skip
    engine.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
