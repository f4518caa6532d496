"""Stand-alone timing of the Laplacian matvec (scs_normalized_matvec_dev) over a range of sizes.

    python tools/matvec_bench.py [m ...]                 (one process: the library's own choice of threads per row)
    SCS_MATVEC_GROUP=64 python tools/matvec_bench.py ... (tuning: force 32 / 64 / 128 / 256 threads per row)

For every m: W symmetric, zero diagonal, density 0.8, entries U(0, 10) (seed 7); 20 launches
back to back on the engine's stream timed with CUDA events, once with W warm in L2 (what a Lanczos run sees when
8 m^2 bytes fit the 126 MB L2) and once with L2 flushed before every launch.  Prints GB/s of the algorithmic bytes
8 m^2 + 24 m."""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from spectralclustersupertree_b200 import _lib  # noqa: E402
from spectralclustersupertree_b200.engine import Engine  # noqa: E402


def main() -> None:
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 2536, 3000, 4000, 4700, 5069, 6000, 7605, 8765]
    engine = Engine(0)
    lib = _lib.load()
    rows = []
    for m in sizes:
        rng = np.random.RandomState(7)
        W = np.triu(rng.uniform(0, 10, (m, m)) * (rng.random_sample((m, m)) < 0.8), 1)
        W = W + W.T
        isd = 1.0 / np.sqrt(W.sum(axis=1))
        x = rng.uniform(-1, 1, m)
        d_W, d_s, d_x, d_y = engine.to_device(W), engine.to_device(isd), engine.to_device(x), engine.alloc(8 * m)
        expect = isd * (W @ (isd * x))
        lib.scs_normalized_matvec_dev(engine.handle, m, d_W, d_s, d_x, d_y)
        got = engine.to_host(d_y, (m,), np.float64)
        assert np.allclose(got, expect, rtol=1e-12, atol=1e-14), m
        nbytes = 8.0 * m * m + 24.0 * m
        for _ in range(5):
            lib.scs_normalized_matvec_dev(engine.handle, m, d_W, d_s, d_x, d_y)
        engine.timer_start()
        for _ in range(20):
            lib.scs_normalized_matvec_dev(engine.handle, m, d_W, d_s, d_x, d_y)
        warm_us = engine.timer_stop() * 1e3 / 20
        cold = []
        for _ in range(8):
            engine.flush_l2()
            engine.timer_start()
            lib.scs_normalized_matvec_dev(engine.handle, m, d_W, d_s, d_x, d_y)
            cold.append(engine.timer_stop() * 1e3)
        cold_us = float(np.median(cold))
        rows.append({"m": m, "warm_us": round(warm_us, 2), "warm_GBps": round(nbytes / warm_us / 1e3),
                     "cold_us": round(cold_us, 2), "cold_GBps": round(nbytes / cold_us / 1e3)})
        for d in (d_W, d_s, d_x, d_y):
            engine.free(d)
    print(json.dumps({"group": os.environ.get("SCS_MATVEC_GROUP", "auto"), "rows": rows}))
    engine.close()


if __name__ == "__main__":
    main()
