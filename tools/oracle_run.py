"""TEST / MEASUREMENT INFRASTRUCTURE -- the CPU oracle's whole recursion on a named workload.

    python tools/oracle_run.py c4 [--trace tests/golden/ctrace_c4.json.gz] [--no-steer] [--seeds 12]

Runs ``oracle.scs_oracle.construct_supertree`` (the CPU restatement of the reference's
``construct_supertree``, ref: /root/reference/src/sc_supertree/scs.py:18-174; C graph build,
numpy components / contraction, sklearn ``SpectralClustering`` exactly as the reference calls it)
on one of bench.py's synthetic workloads, from ``PhyloNode`` objects, and

* prints one JSON line with the wall-clock seconds of the reference-equivalent work, split per stage
  (this is what ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` report: measured, never
  extrapolated), and
* optionally writes a *compact trace* -- one record per recursion node -- that the GPU parity tests
  compare against node by node at full size (tests/test_gpu_fullsize.py).

Compact trace record (vertex ids = global taxon ids = ranks in sorted(all tip names)):
  key   blake2b-64 of the node's ascending int32 taxon ids
  n     taxa at the node;  nc  components of the proper cluster graph
  m     contracted size (spectral nodes only)
  part  hash of the canonical partition the recursion continued with (side of the first vertex = 0;
        components numbered by first appearance)
  nat   the partition sklearn returned on the RandomState(0) stream, when it is not ``part``
  eig   [lambda_2, lambda_3] of the normalised Laplacian of the contracted graph
  margin  distance of the closest vertex to the 2-means boundary / range of the Fiedler coordinate
  km    only where the 1-D 2-means of the Fiedler coordinate has several Lloyd-stable splits:
        {"stable": count, "seen": [hashes of the partitions sklearn's k-means returned over --seeds RNG seeds],
         "opt": hash of the global optimum}

Steering (default on): at a node where sklearn's k-means outcome depends on its RNG seed (several
distinct partitions over the seeds) and the global optimum of the 1-D 2-means is one of the outcomes
observed, the recursion continues with that optimum -- a possible run of the reference, and the one a
deterministic exact 2-means reproduces.  ``nat`` keeps what the seed-0 stream returned.
"""

from __future__ import annotations

import argparse
import gzip
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GAP_TIE = 1e-7
MARGIN_TIE = 1e-9


def key_of(ids: np.ndarray) -> str:
    return hashlib.blake2b(np.ascontiguousarray(ids, dtype="<i4").tobytes(), digest_size=8).hexdigest()


def canonical_sides(side: np.ndarray) -> np.ndarray:
    """Labels renumbered by first appearance (a bipartition: side of the first vertex = 0)."""
    values, first, inverse = np.unique(np.asarray(side), return_index=True, return_inverse=True)
    rank = np.empty(len(values), dtype=np.int32)
    rank[np.argsort(first)] = np.arange(len(values), dtype=np.int32)
    return rank[inverse]


def part_hash(labels: np.ndarray) -> str:
    return hashlib.blake2b(canonical_sides(labels).astype("<i4").tobytes(), digest_size=8).hexdigest()


def small_eigs(Wc: np.ndarray):
    """(lambda_2, lambda_3, Fiedler coordinate u_1) of the contracted graph."""
    from oracle import scs_oracle

    m = Wc.shape[0]
    if m <= 1500:
        vals, emb = scs_oracle.normalized_affinity_eigs(Wc, 3)
        lam3 = float(vals[2]) if len(vals) > 2 else float("nan")
        return float(vals[1]), lam3, emb[:, 1]
    from scipy.sparse.linalg import eigsh

    A = np.array(Wc, dtype=np.float64)
    np.fill_diagonal(A, 0.0)
    d = A.sum(axis=0)
    s = np.where(d == 0, 1.0, np.sqrt(d))
    N = A / s[:, None] / s[None, :]
    vals, vecs = eigsh(N, k=3, which="LA", tol=1e-13, v0=np.ones(m))
    order = np.argsort(-vals)
    vals, vecs = vals[order], vecs[:, order]
    return float(1.0 - vals[1]), float(1.0 - vals[2]), vecs[:, 1] / s


def trace_recursion(trees, weights, weighting: str, names, steer: bool, seeds: int, want_trace: bool = True,
                    contract_edges: bool = True):
    """The oracle's whole recursion on ``trees`` (PhyloNode objects; ``names`` = sorted taxon names, global taxon id
    = index).  Returns ``(summary, compact trace records, supertree)``; the records are empty without ``want_trace``."""
    from oracle import scs_oracle

    gid = {name: i for i, name in enumerate(names)}
    scs_oracle._c_lib()
    records: list[dict] = []

    def hook(record, Wc, side, groups):
        # side: sklearn's labels per contracted vertex; groups: vertex -> contracted vertex
        ids = np.fromiter((gid[x] for x in record["names"]), dtype=np.int32, count=len(record["names"]))
        order = np.argsort(ids)
        m = Wc.shape[0]
        info = {"m": int(m)}
        natural = np.asarray(side)[groups][order]
        chosen = natural
        replacement = None
        if m > 2:
            lam2, lam3, u1 = small_eigs(Wc)
            info["eig"] = [lam2, lam3]
            s0 = np.asarray(side)
            if 0 < s0.sum() < m:
                mid = 0.5 * (u1[s0 == 0].mean() + u1[s0 == 1].mean())
                info["margin"] = float(np.abs(u1 - mid).min() / max(np.ptp(u1), 1e-300))
            srt = np.argsort(u1, kind="stable")
            s = u1[srt]
            c = s - s.mean()
            prefix = np.cumsum(c)[:-1]
            cnt = np.arange(1, m)
            score = prefix**2 / cnt + (c.sum() - prefix) ** 2 / (m - cnt)
            csum = np.cumsum(s)
            lo_mean = csum[:-1] / cnt
            hi_mean = (csum[-1] - csum[:-1]) / (m - cnt)
            mid = 0.5 * (lo_mean + hi_mean)
            stable = int(np.count_nonzero((s[:-1] < mid) & (mid < s[1:])))
            if stable > 1:
                from sklearn.cluster import k_means
                from sklearn.manifold import spectral_embedding

                best = int(np.argmax(score)) + 1
                opt_side = np.zeros(m, dtype=np.int32)
                opt_side[srt[best:]] = 1
                opt = opt_side[groups][order]
                seen = {}
                maps = spectral_embedding(np.array(Wc, dtype=np.float64), n_components=2, eigen_solver=None,
                                          random_state=np.random.RandomState(0), drop_first=False)
                for sd in range(seeds):
                    _, lab, _ = k_means(maps, 2, random_state=np.random.RandomState(1000 + sd), n_init=10)
                    full = np.asarray(lab)[groups][order]
                    seen[part_hash(full)] = full
                seen[part_hash(natural)] = natural
                info["km"] = {"stable": stable, "seen": sorted(seen), "opt": part_hash(opt)}
                if steer and len(seen) > 1 and part_hash(opt) in seen and part_hash(opt) != part_hash(natural):
                    chosen = opt
                    info["nat"] = part_hash(natural)
                    replacement = opt_side
        info["part"] = part_hash(chosen)
        info["sizes"] = [int((chosen == chosen[0]).sum()), int((chosen != chosen[0]).sum())]
        record["_info"] = info
        return replacement

    timers: dict = {}
    trace: list = []
    t0 = time.perf_counter()
    tree = scs_oracle.construct_supertree(
        trees, weights, weighting, random_state=np.random.RandomState(0), use_c=True, contract_edges=contract_edges,
        trace=trace, timers=timers, node_hook=hook if want_trace else None,
    )  # fmt: skip
    wall = time.perf_counter() - t0
    hook_s = timers.pop("hook", 0.0)
    total = wall - hook_s
    out = {
        "seconds": total, "stages": timers, "other_s": total - sum(timers.values()), "hook_s": hook_s,
        "recursion_nodes": len(trace), "spectral_nodes": sum(1 for r in trace if "partition" in r),
        "cores": os.cpu_count(), "steered": bool(steer and want_trace),
    }  # fmt: skip
    if want_trace:
        for rec in trace:
            ids = np.sort(np.fromiter((gid[x] for x in rec["names"]), dtype=np.int32, count=len(rec["names"])))
            item = {"key": key_of(ids), "n": len(ids), "nc": rec["n_components"]}
            if "_info" in rec:
                item.update(rec["_info"])
            records.append(item)
    return out, records, tree


def write_trace(trace_path: Path, payload: dict, tree) -> None:
    from spectralclustersupertree_b200.tree import make_tree

    payload["supertree"] = tree.get_newick()
    opener = gzip.open if str(trace_path).endswith(".gz") else open
    with opener(trace_path, "wt") as fh:
        json.dump(payload, fh)
    # the written supertree must parse back to the same clades
    assert make_tree(payload["supertree"]).clade_sets() == tree.clade_sets()


def run(workload: str, trace_path: Path | None, steer: bool, seeds: int) -> dict:
    from bench import WORKLOADS, describe
    from spectralclustersupertree_b200.synthetic import make_problem

    n, t, weighting, seed, tw = WORKLOADS[workload]
    prob = make_problem(n, t, weighting, seed, tree_weights=tw)
    t0 = time.perf_counter()
    trees = prob.phylonodes()
    parse_s = time.perf_counter() - t0
    weights = [1.0] * len(trees) if prob.weights is None else list(prob.weights)
    names = prob.names()
    want_trace = trace_path is not None
    summary, records, tree = trace_recursion(trees, weights, weighting, names, steer, seeds, want_trace)
    out = {"workload": workload, "describe": describe(workload), "phylonode_build_s": parse_s, **summary}
    if want_trace:
        payload = dict(out)
        payload["names"] = len(names)
        payload["nodes"] = records
        payload["seeds"] = seeds
        write_trace(trace_path, payload, tree)
    return out


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("workload")
    parser.add_argument("--trace", type=Path, default=None)
    parser.add_argument("--no-steer", action="store_true")
    parser.add_argument("--seeds", type=int, default=12)
    args = parser.parse_args()
    print(json.dumps(run(args.workload, args.trace, not args.no_steer, args.seeds)), flush=True)


if __name__ == "__main__":
    main()
