"""How many passes over W does a block eigensolver save?  (CPU experiment behind DESIGN.md's tensor-pipe paragraph.)

    python tools/block_lanczos_experiment.py [workload]     # default c4; ~1 min of CPU, 1.3 GB of memory

Builds the top-level graph of the workload with the C oracle, takes its largest connected component (C4: m = 8 765, the
largest spectral node of the job), and runs block Lanczos with full re-orthogonalisation on the deflated operator
N = D^-1/2 W D^-1/2 for block sizes 1, 2, 4, 8 until the Fiedler pair's true residual is <= 1e-12 (the GPU solver's
stopping rule).  Prints the number of block steps = passes over W.  Test infrastructure: imports oracle/.
"""

from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from oracle import scs_oracle  # noqa: E402


def largest_component_matrix(workload: str) -> np.ndarray:
    a = bench.make_workload(workload)
    n = len(a["names"])
    roots, child_ptr, child_idx, tip_taxon, own = bench.oracle_children_csr(a)
    W, C, _ = scs_oracle.pcg_dense_c_arrays(n, roots, child_ptr, child_idx, tip_taxon, own, a["weights"], a["weighting"])
    label = scs_oracle.graph_components(C > 0)
    values, counts = np.unique(label, return_counts=True)
    idx = np.flatnonzero(label == values[np.argmax(counts)])
    return np.ascontiguousarray(W[np.ix_(idx, idx)])


def block_lanczos(W: np.ndarray, b: int, tol: float = 1e-12, max_steps: int = 100):
    m = W.shape[0]
    d = W.sum(1)
    isd = 1.0 / np.sqrt(d)
    q0 = np.sqrt(d)
    q0 /= np.linalg.norm(q0)
    rng = np.random.RandomState(0)

    def op(X):  # one pass over W
        Y = isd[:, None] * (W @ (isd[:, None] * X))
        return Y - np.outer(q0, q0 @ Y)

    X = rng.uniform(-1, 1, (m, b))
    X -= np.outer(q0, q0 @ X)
    basis = [np.linalg.qr(X)[0]]
    NV = None
    for step in range(1, max_steps + 1):
        Y = op(basis[-1])
        V = np.concatenate(basis, 1)
        NV = Y if NV is None else np.concatenate([NV, Y], 1)
        H = V.T @ NV
        w, S = np.linalg.eigh((H + H.T) / 2)
        y = V @ S[:, -1]
        residual = np.linalg.norm(NV @ S[:, -1] - w[-1] * y)
        if residual <= tol:
            return step, 1.0 - w[-1], 1.0 - w[-2]
        for _ in range(2):
            Y = Y - V @ (V.T @ Y)
        basis.append(np.linalg.qr(Y)[0])
    return None


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    W = largest_component_matrix(workload)
    print(f"{workload}: largest component of the top-level graph, m = {W.shape[0]}")
    for b in (1, 2, 4, 8):
        t0 = time.perf_counter()
        steps, lam2, lam3 = block_lanczos(W, b)
        print(f"block size {b}: {steps} passes over W (lambda2 {lam2:.12f}, lambda3 {lam3:.12f}; {time.perf_counter() - t0:.1f} s)")


if __name__ == "__main__":
    main()
