"""Benchmark of the supertree hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4] [--impl reference]

One *step* is one complete ``construct_supertree`` job on the named synthetic workload (default
``c4``: 10 000 taxa x 1 000 source trees, depth weighting -- the configuration BASELINE.json's metric
is quoted on).  Three clocks are reported in one JSON line:

* ``value``  -- the hot path alone with inputs resident in HBM: every recursion node of the job
  (graph build, components, contraction, spectral split) replayed from leaf tours that were uploaded
  before the timed region; CUDA events on the engine's stream.
* ``e2e``    -- the same job through the public host-buffer path (``supertree_of_forest`` over the
  C ABI): flat source trees in host memory in, supertree out; host recursion, tree restriction,
  H2D of every node's tours and D2H of every node's partition inside the timed region.
* ``roofline`` -- the Laplacian matvec (the kernel BASELINE.json's metric names), timed per launch
  with CUDA events inside the timed steps, against the measured HBM peak.

``cpu_baseline`` / ``--impl reference`` time the CPU oracle port of the reference (oracle/: the
reference is pure Python and needs cogent3, which is not installable here -- DESIGN.md) on a bounded
sample: the job's top-level recursion node.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "construct_supertree wall-time at 10k taxa/1k trees"

# name -> (taxa, trees, weighting, seed, tree weights)   (BASELINE.json configs; seed = 1000 * index)
WORKLOADS = {
    "c1": (100, 30, "depth", 1000, False),
    "c2": (500, 50, "branch", 2000, False),
    "c3": (1000, 100, "branch", 3000, True),
    "c4": (10000, 1000, "depth", 4000, False),
}


def describe(workload: str) -> str:
    n, t, weighting, seed, tw = WORKLOADS[workload]
    extra = ", tree weights U[0.5,2]" if tw else ""
    return (
        f"{workload}: {n} taxa x {t} source trees (birth-death model tree, SMIDGen-style subsets, 5% NNI noise), "
        f"pcg_weighting={weighting}{extra}, seed {seed}"
    )


def make_workload(workload: str) -> dict:
    from spectralclustersupertree_b200.synthetic import make_problem

    n, t, weighting, seed, tw = WORKLOADS[workload]
    prob = make_problem(n, t, weighting, seed, tree_weights=tw)
    arrays = prob.forest_arrays()
    arrays["weighting"] = weighting
    return arrays


def measured_peak_gbs() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.is_file():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device: int) -> None:
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            cells = [c.strip() for c in line.split(",")]
            if len(cells) < 7:
                continue
            try:
                sm.append(float(cells[0]))
                sm_max.append(float(cells[1]))
            except ValueError:
                continue
            for name, cell in zip(names, cells[3:7], strict=True):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {
            "sm_mhz": statistics.median(sm),
            "sm_max_mhz": max(sm_max),
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ---------------------------------------------------------------------------------------------
# the CPU oracle on a bounded sample (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
def oracle_children_csr(arrays: dict):
    """The flat pre-order forest as the child lists oracle/pcg_oracle.c reads (vectorised)."""
    offsets = arrays["node_offsets"]
    parent = arrays["parent"].astype(np.int64)
    tree_of = np.repeat(np.arange(len(offsets) - 1), np.diff(offsets))
    base = offsets[:-1][tree_of]
    is_root = parent < 0
    gparent = np.where(is_root, -1, parent + base)
    order = np.argsort(gparent, kind="stable")  # children of a node, in pre-order = child order
    order = order[gparent[order] >= 0]
    counts = np.bincount(gparent[order], minlength=len(parent))
    child_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    child_idx = order.astype(np.int64)
    weighting = arrays["weighting"]
    if weighting == "branch":
        own = np.where(np.isnan(arrays["length"]), 1.0, arrays["length"])
    elif weighting == "bootstrap":
        own = arrays["support"].copy()
    else:
        own = np.zeros(len(parent))
    return offsets[:-1].astype(np.int64), child_ptr, child_idx, arrays["taxon"].astype(np.int32), own.astype(np.float64)


def oracle_top_node(arrays: dict) -> dict:
    """Time the CPU oracle on the job's top-level recursion node (ref: scs.py:110-134)."""
    from oracle import scs_oracle

    n = len(arrays["names"])
    roots, child_ptr, child_idx, tip_taxon, own = oracle_children_csr(arrays)
    scs_oracle._c_lib()  # build / load outside the timed region
    t0 = time.perf_counter()
    W, C, occ = scs_oracle.pcg_dense_c_arrays(
        n, roots, child_ptr, child_idx, tip_taxon, own, arrays["weights"], arrays["weighting"]
    )
    t1 = time.perf_counter()
    label = scs_oracle.graph_components(C > 0)
    n_comp = len(np.unique(label))
    t2 = time.perf_counter()
    spectral_n = 0
    if n_comp == 1:
        _, Wc, _ = scs_oracle.contract_dense(W, C, occ)
        spectral_n = Wc.shape[0]
        scs_oracle.spectral_bipartition(Wc, np.random.RandomState(0))
    else:
        # the graph is disconnected at the top: the spectral stage first runs one level down, on
        # the largest component; time sklearn on that component's graph so the sample covers it
        sizes = np.bincount(label)
        big = np.flatnonzero(label == np.argmax(sizes))
        if len(big) >= 3:
            sub = np.ix_(big, big)
            _, Wc, _ = scs_oracle.contract_dense(W[sub], C[sub], occ[big])
            if Wc.shape[0] >= 2 and len(np.unique(scs_oracle.graph_components(Wc > 0))) == 1:
                spectral_n = Wc.shape[0]
                scs_oracle.spectral_bipartition(Wc, np.random.RandomState(0))
    t3 = time.perf_counter()
    return {
        "seconds": t3 - t0,
        "pcg_s": t1 - t0,
        "components_s": t2 - t1,
        "spectral_s": t3 - t2,
        "n": n,
        "n_components": n_comp,
        "spectral_n": int(spectral_n),
    }


def workload_ratio(workload: str) -> dict | None:
    """Whole-job / top-level-node work ratio recorded by a GPU run (profiles/workloads.json)."""
    path = ROOT / "profiles" / "workloads.json"
    if not path.is_file():
        return None
    return json.loads(path.read_text()).get(workload)


def cpu_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def reference_line(args, arrays: dict) -> dict:
    """``--impl reference``: the CPU oracle port, each step = the bounded sample."""
    times = []
    detail = {}
    for i in range(args.warmup + args.steps):
        detail = oracle_top_node(arrays)
        if i >= args.warmup:
            times.append(detail["seconds"])
    sample_s = statistics.mean(times)
    ratio = workload_ratio(args.workload)
    scale = ratio["pair_visits_total"] / ratio["pair_visits_top"] if ratio else 1.0
    value = sample_s * scale
    sample = (
        f"top-level recursion node of {args.workload} (C oracle graph build over all trees {detail['pcg_s']:.1f} s, "
        f"components {detail['components_s']:.1f} s, contraction + sklearn SpectralClustering on "
        f"{detail['spectral_n']} vertices {detail['spectral_s']:.1f} s) = {sample_s:.1f} s measured; "
        f"whole job extrapolated x{scale:.2f} by leaf-pair updates over all recursion nodes"
    )
    cores = cpu_threads()
    return {
        "metric": METRIC, "value": value, "unit": "s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sample_s * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload)},
        "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------
class Replay:
    """Every recursion node's leaf tours, resident in HBM, for the device-only timed pass."""

    def __init__(self, engine, nodes: list) -> None:
        self.engine = engine
        self.nodes = []
        self.pair_visits = [t.pair_updates() for t, _ in nodes]
        cat = lambda field, dtype: np.concatenate([getattr(t, field) for t, _ in nodes]).astype(dtype)  # noqa: E731
        host = {
            "leaf_offsets": cat("leaf_offsets", np.int64), "leaf_taxon": cat("leaf_taxon", np.int32),
            "adj_depth": cat("adj_depth", np.int32), "adj_val": cat("adj_val", np.float64),
            "root_depth": cat("root_depth", np.int32), "tree_weight": cat("tree_weight", np.float64),
        }  # fmt: skip
        self.bytes = sum(a.nbytes for a in host.values())
        self.dev = {k: engine.to_device(v) for k, v in host.items()}
        size = {k: v.itemsize for k, v in host.items()}
        pos = dict.fromkeys(host, 0)
        n_max = 1
        for tours, seed in nodes:
            T, L = tours.num_trees, tours.num_leaves
            entry = {"n": tours.n, "T": T, "L": L, "seed": seed}
            for key, count in (("leaf_offsets", T + 1), ("leaf_taxon", L), ("adj_depth", L), ("adj_val", L),
                               ("root_depth", T), ("tree_weight", T)):  # fmt: skip
                entry[key] = self.dev[key] + pos[key] * size[key]
                pos[key] += count
            self.nodes.append(entry)
            n_max = max(n_max, tours.n)
        self.part = engine.alloc(4 * n_max)

    def run(self, contract_edges: bool = True) -> None:
        for entry in self.nodes:
            self.engine.node_split_dev(entry, self.part, contract_edges=contract_edges, seed=entry["seed"])

    def close(self) -> None:
        for d in self.dev.values():
            self.engine.free(d)
        self.engine.free(self.part)


def gpu_line(args, arrays: dict) -> dict:
    from spectralclustersupertree_b200.engine import Engine, Forest
    from spectralclustersupertree_b200.scs import supertree_of_forest

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    engine = Engine(local_rank)
    weighting = arrays["weighting"]

    def new_forest():
        return Forest.from_arrays(arrays["node_offsets"], arrays["parent"], arrays["length"], arrays["support"],
                                  arrays["taxon"], arrays["weights"], arrays["names"])  # fmt: skip

    # recording pass (untimed): the recursion nodes of this job, for the device-resident replay
    recorded: list = []
    trace: list = []
    tree = supertree_of_forest(new_forest(), weighting, engine=engine, trace=trace,
                               node_hook=lambda f, seed: recorded.append((f.tours(weighting), seed)))  # fmt: skip
    n_tips = len(tree.get_tip_names())
    replay = Replay(engine, recorded)
    spectral = [r for r in trace if r["n_components"] == 1]
    job = {
        "recursion_nodes": len(trace),
        "spectral_nodes": len(spectral),
        "largest_spectral_m": max((r["contracted_size"] for r in spectral), default=0),
        "matvecs": sum(r["stats"]["matvecs"] for r in spectral),
        "tie_nodes": sum(1 for r in spectral if r["stats"]["tie_flag"]),
        "pair_visits_total": int(sum(replay.pair_visits)),
        "pair_visits_top": int(replay.pair_visits[0]) if replay.pair_visits else 0,
        "tour_bytes": int(replay.bytes),
        "supertree_tips": n_tips,
    }

    def barrier():
        engine.synchronize()
        if dist is not None:
            dist.barrier()

    def e2e_step():
        return supertree_of_forest(new_forest(), weighting, engine=engine)

    for _ in range(args.warmup):
        e2e_step()
        replay.run()
    engine.synchronize()

    sampler = ClockSampler(local_rank)
    # ---- timed: device-resident replay (value) with per-launch matvec / row-kernel timing ----------
    engine.profile(True)
    launches_before = engine.launch_count
    dev_ms = []
    for _ in range(args.steps):
        engine.flush_l2()
        barrier()
        engine.timer_start()
        replay.run()
        dev_ms.append(engine.timer_stop())
        barrier()
    gpu_launches = engine.launch_count - launches_before
    engine.profile(False)
    matvec = engine.profile_read(0)
    rows = engine.profile_read(1)
    # ---- timed: end to end from host buffers (e2e) -----------------------------------------------
    h2d0, d2h0 = engine.io_bytes()
    e2e_s = []
    for _ in range(args.steps):
        engine.flush_l2()
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        engine.synchronize()
        e2e_s.append(time.perf_counter() - t0)
        barrier()
    h2d1, d2h1 = engine.io_bytes()
    clocks = sampler.stop()

    value_s = statistics.mean(dev_ms) / 1e3
    e2e_value = statistics.mean(e2e_s)
    if dist is not None:
        import torch

        both = torch.tensor([value_s, e2e_value], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        value_s, e2e_value = both.tolist()

    peak, peak_src = measured_peak_gbs()
    roofline = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "kernel": "matvec_row_per_cta (y = D^-1/2 W D^-1/2 x, fp64, m >= 2048)", "peak_source": peak_src}  # fmt: skip
    if matvec["launches"] and matvec["ms"] > 0:
        achieved = matvec["bytes"] / (matvec["ms"] * 1e-3) / 1e9
        roofline.update({
            "achieved": achieved, "frac": achieved / peak, "launches": matvec["launches"],
            "avg_launch_us": 1e3 * matvec["ms"] / matvec["launches"],
            "bytes_per_launch": matvec["bytes"] / matvec["launches"],
            "share_of_step": matvec["ms"] / sum(dev_ms),
        })  # fmt: skip
    roofline_rows = None
    if rows["launches"] and rows["ms"] > 0:
        roofline_rows = {
            "kernel": "pcg_rows_kernel (leaf-pair LCA weighting -> W rows, adjacency bits, degree)",
            "bound": "shared-memory / issue (not HBM): leaf-pair visits per second",
            "pair_visits_per_s": rows["units"] / (rows["ms"] * 1e-3),
            "write_GBps": rows["bytes"] / (rows["ms"] * 1e-3) / 1e9,
            "launches": rows["launches"],
            "share_of_step": rows["ms"] / sum(dev_ms),
        }
    line = {
        "metric": METRIC, "value": value_s, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": value_s * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": describe(args.workload),
            "l2": "flushed between timed steps (256 MB write); the top-level W (0.8 GB) exceeds L2 by itself",
            "value_is": "all recursion nodes replayed from leaf tours resident in HBM (CUDA events)",
            "e2e_is": "supertree_of_forest over the C ABI from flat host arrays (wall clock)",
            "job": job,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "s", "h2d_bytes_per_step": (h2d1 - h2d0) // args.steps,
                "d2h_bytes_per_step": (d2h1 - d2h0) // args.steps},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline,
        "roofline_pcg_rows": roofline_rows,
    }  # fmt: skip
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        detail = oracle_top_node(arrays)
        scale = job["pair_visits_total"] / max(job["pair_visits_top"], 1)
        line["cpu_baseline"] = {
            "value": detail["seconds"] * scale, "unit": "s", "cores": cpu_threads(), "kind": "port",
            "sample": (
                f"top-level recursion node of {args.workload} on the CPU oracle (C graph build {detail['pcg_s']:.1f} s, "
                f"components {detail['components_s']:.1f} s, contraction + sklearn SpectralClustering on "
                f"{detail['spectral_n']} vertices {detail['spectral_s']:.1f} s) = {detail['seconds']:.1f} s measured; "
                f"whole job extrapolated x{scale:.2f} by leaf-pair updates over all recursion nodes"
            ),
        }  # fmt: skip
        out = ROOT / "gpurun_out"
        if out.is_dir():
            (out / f"workload_{args.workload}.json").write_text(json.dumps({args.workload: job}, indent=1) + "\n")
    replay.close()
    engine.close()
    if dist is not None:
        dist.destroy_process_group()
    return line if rank == 0 else {}


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=3)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(reference_line(args, make_workload(args.workload))), flush=True)
        return
    line = gpu_line(args, make_workload(args.workload))
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
