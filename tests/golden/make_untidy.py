"""TEST INFRASTRUCTURE -- golden cases with *untidy* source trees, traced by the CPU oracle.

    PYTHONHASHSEED=0 python tests/golden/make_untidy.py

The synthetic workloads and the reference's fixtures consist of tidy trees: every internal node branches, every
non-root node of a branch-weighted tree has a length.  The reference accepts more than that -- it walks whatever
``PhyloNode`` it is given (ref: /root/reference/src/sc_supertree/scs.py:569-579, 624-631): internal nodes with a
single child (each counts one level of depth and adds its length), polytomies, nodes without a length (worth 1 under
``branch``, ref: scs.py:560-562), a root with a single child (one "side" that holds every taxon), trees of two tips
and lone tips.  The restriction of such trees to a component (``get_sub_tree``, ref: scs.py:444-453) merges the unary
nodes away, so the first recursion level sees them and the deeper ones see their merged lengths.

This script makes three seeded cases of that kind (SURVEY.md section 8d's recipe, then perturbed), runs the oracle's
whole recursion on them (``tools/oracle_run.py:trace_recursion``: C graph build, numpy components / contraction,
sklearn ``SpectralClustering`` as the reference calls it, ``tree.py`` restriction) and writes
``tests/golden/ctrace_untidy_<case>.json.gz``: the Newick lines, the tree weights and one compact record per
recursion node.  ``tests/test_gpu_fullsize.py`` builds the same supertrees on the GPU from those lines -- device-
resident forest, host forest, and the native Newick parser -- and compares node by node.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
for path in (ROOT, ROOT / "tools", ROOT / "tests"):
    if str(path) not in sys.path:
        sys.path.insert(0, str(path))

# name -> (taxa, trees, weighting, seed, tree weights)
CASES = {
    "branch": (150, 40, "branch", 9100, True),
    "bootstrap": (120, 30, "bootstrap", 9200, False),
    "depth": (200, 40, "depth", 9300, False),
    # the same kind of trees with contract_edges=False (ref: scs.py:23,124-133): the spectral step on the whole graph
    "nocontract": (400, 50, "depth", 9500, False),
    # 20 bushy trees + 12 caterpillars of ~750 tips each (tip order = model order with noise): leaf tours more than
    # 1024 levels deep at the top of the recursion, long chains of unary nodes to merge below it
    "caterpillar": (1500, 20, "branch", 9600, False),
}
NO_CONTRACTION = {"nocontract"}


def caterpillars(n: int, count: int, rng: np.random.RandomState) -> list[str]:
    """``count`` caterpillar trees ((((a,b),c),d),...) on random halves of t0..t{n-1}, with branch lengths."""
    lines = []
    for _ in range(count):
        picked = np.flatnonzero(rng.random_sample(n) < 0.5)
        order = picked[np.argsort(picked + rng.normal(0.0, 0.02 * n, size=len(picked)), kind="stable")]
        text = f"(t{order[0]}:{rng.uniform(0.1, 1.0)!r},t{order[1]}:{rng.uniform(0.1, 1.0)!r})"
        for x in order[2:]:
            text = f"({text}:{rng.uniform(0.1, 1.0)!r},t{x}:{rng.uniform(0.1, 1.0)!r})"
        lines.append(text + ";")
    return lines


def make_untidy(topo, weighting: str, rng: np.random.RandomState) -> None:
    """Perturb one source tree in place: polytomies, unary chains, missing lengths, a unary root."""
    parent = {c: x for x in topo.preorder() for c in topo.children[x]}

    def support():
        return float(rng.randint(50, 101)) if weighting == "bootstrap" else None

    # polytomies: an internal non-root node gives its children to its parent
    for x in topo.preorder():
        if x != topo.root and topo.children[x] and rng.random_sample() < 0.08:
            p = parent[x]
            at = topo.children[p].index(x)
            topo.children[p][at : at + 1] = topo.children[x]
            for c in topo.children[x]:
                parent[c] = p
    # unary nodes: chains of one to three on the branch above a node, the branch's length shared out
    for x in topo.preorder():
        if x == topo.root or rng.random_sample() >= 0.15:
            continue
        for _ in range(1 + int(rng.random_sample() < 0.3) + int(rng.random_sample() < 0.3)):
            p = parent[x]
            u = topo.add(None, None, support())
            if topo.length[x] is not None:
                share = float(rng.uniform(0.2, 0.8))
                topo.length[u] = topo.length[x] * share
                topo.length[x] = topo.length[x] * (1.0 - share)
            topo.children[p][topo.children[p].index(x)] = u
            topo.children[u] = [x]
            parent[u], parent[x] = p, u
    # a root with a single child: the old root becomes the only side
    if rng.random_sample() < 0.25:
        old = topo.root
        top = topo.add(None, None, None)
        topo.children[top] = [old]
        topo.root = top
        topo.support[old] = support()
        if weighting == "branch" and rng.random_sample() < 0.5:
            topo.length[old] = float(rng.uniform(0.1, 2.0))
    # missing lengths (worth 1 under branch weighting)
    if weighting == "branch":
        for x in topo.preorder():
            if x != topo.root and rng.random_sample() < 0.1:
                topo.length[x] = None


def untidy_lines(case: str) -> tuple[list[str], list[float], str]:
    from spectralclustersupertree_b200.synthetic import make_problem

    n, trees, weighting, seed, tree_weights = CASES[case]
    problem = make_problem(n, trees, weighting, seed, tree_weights=tree_weights)
    rng = np.random.RandomState(seed + 1)
    for topo in problem.sources:
        make_untidy(topo, weighting, rng)
    lines = problem.newick_lines()
    weights = [1.0] * len(lines) if problem.weights is None else list(problem.weights)
    # trees that contribute occurrences but no proper cluster, and one that contributes nothing
    tips = sorted({name for s in problem.sources for name in s.tip_names()})
    inner = "75" if weighting == "bootstrap" else ""  # the only side of a unary root is weighted like any other node
    lines += [f"({tips[0]},{tips[-1]});", f"{tips[1]};", f"(({tips[2]},{tips[3]}){inner});"]
    weights += [1.5, 1.0, 2.0]
    if case == "caterpillar":
        extra = caterpillars(n, 12, rng)
        lines += extra
        weights += [1.0] * len(extra)
    return lines, weights, weighting


def main() -> None:
    from helpers import parse
    from oracle_run import trace_recursion, write_trace

    for case in sys.argv[1:] or CASES:
        lines, weights, weighting = untidy_lines(case)
        trees = parse(lines)
        names = sorted({x for t in trees for x in t.get_tip_names()})
        unary = sum(1 for t in trees for node in t.preorder() if len(node.children) == 1)
        wide = sum(1 for t in trees for node in t.preorder() if len(node.children) > 2)
        missing = sum(1 for t in trees for node in t.preorder(include_self=False) if node.length is None)
        contract = case not in NO_CONTRACTION
        summary, records, tree = trace_recursion(trees, weights, weighting, names, steer=True, seeds=12,
                                                 contract_edges=contract)  # fmt: skip
        payload = {
            "case": case, "weighting": weighting, "contract_edges": contract, "lines": lines, "weights": weights, "names": len(names),
            "unary_nodes": unary, "polytomies": wide, "nodes_without_length": missing, "nodes": records, "seeds": 12,
            **summary,
        }  # fmt: skip
        write_trace(Path(__file__).resolve().parent / f"ctrace_untidy_{case}.json.gz", payload, tree)
        print(case, {k: payload[k] for k in ("names", "unary_nodes", "polytomies", "nodes_without_length",
                                             "recursion_nodes", "spectral_nodes", "seconds")})  # fmt: skip


if __name__ == "__main__":
    main()
