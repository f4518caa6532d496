"""Debug aid: W row blocks of an in-process 2-rank sharded node against the single-GPU W (run on the GPU box)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_sharded_nodes import Ranks, forest_of  # noqa: E402

from spectralclustersupertree_b200 import _lib  # noqa: E402
from spectralclustersupertree_b200.engine import Engine  # noqa: E402
from spectralclustersupertree_b200.synthetic import make_problem  # noqa: E402


def layout(n_max, world):
    words = (n_max + 31) // 32
    at = 4096
    out = {}

    def take(name, nbytes):
        nonlocal at
        out[name] = at
        at = (at + nbytes + 255) // 256 * 256

    take("vec0", 8 * n_max)
    take("vec1", 8 * n_max)
    take("degree", 8 * n_max)
    take("degree_c", 8 * n_max)
    take("adj", 4 * n_max * words)
    take("max", 4 * n_max * words)
    take("W", 8 * ((n_max + world - 1) // world) * n_max)
    return out


def main():
    n, trees, weighting, world, contract = 900, 80, "one", 2, False
    engine = Engine(0)
    arrays = make_problem(n, trees, weighting, 77 + n, tree_weights=False).forest_arrays()
    forest = forest_of(arrays)
    tours = forest.tours(weighting)
    want_part, want = engine.node_split(tours, contract_edges=contract, seed=5)
    while want.n_components != 1:
        sizes = np.bincount(want_part)
        forest = forest.induce(forest.taxa()[want_part == np.argmax(sizes)])
        tours = forest.tours(weighting)
        want_part, want = engine.node_split(tours, contract_edges=contract, seed=5)
    ref = engine.last_node_buffers()
    m = tours.n
    print("node n", m, "trees", tours.num_trees)
    ranks = Ranks(world, n_max=m, min_n=64)
    ranks.run(lambda r, eng: eng.node_split(tours, contract_edges=contract, seed=5))
    for eng in ranks.engines:
        eng.shard_engage(True)
    got = ranks.run(lambda r, eng: eng.node_split(tours, contract_edges=contract, seed=5))
    lay = layout(m, world)
    rpr = (m + world - 1) // world
    for r, eng in enumerate(ranks.engines):
        eng.synchronize()
        base = eng.shard_window()
        r0, r1 = r * rpr, min(m, (r + 1) * rpr)
        Wb = eng.to_host(base + lay["W"], (r1 - r0, m), np.float64)
        deg = eng.to_host(base + lay["degree"], (m,), np.float64)
        bad = np.argwhere(Wb != ref["W"][r0:r1])
        print("rank", r, "rows", r0, r1, "W mismatches", len(bad), "degree mismatches", int((deg != ref["degree"]).sum()))
        for a, c in bad[:12]:
            print("   row", a + r0, "col", c, "got", Wb[a, c], "want", ref["W"][a + r0, c], "d", (c - a - r0) % m)
        words = (m + 31) // 32
        adj = eng.to_host(base + lay["adj"], (m, words), np.uint32)
        print("   adj mismatching words", int((adj != ref["adj_bits"]).sum()))
    print("eig", [s.eig[1] for _, s in got], want.eig[1])
    ranks.close()


if __name__ == "__main__":
    main()
