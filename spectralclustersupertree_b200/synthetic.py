"""Seeded synthetic supertree problems (SURVEY.md section 8d recipe).

A birth-death model tree on ``n`` taxa and ``T`` SMIDGen-style source trees: tree 0 is a
scaffold on a uniform 20 % sample of the taxa, the others are clade-based (a random clade of at
least 5 % of the taxa, each of its tips kept with p = 0.5, capped at 25 % of the taxa, plus one
outgroup tip), each perturbed by a few random nearest-neighbour interchanges so that the source
trees conflict and the spectral stage is actually exercised.

Trees are held as ``Topology`` (child lists over integer node ids) so the 10 000-taxon
configurations never build per-node Python objects unless asked to (``to_phylonode``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .tree import PhyloNode


@dataclass
class Topology:
    """A rooted tree as child lists; node 0.. in creation order, ``root`` names the root."""

    children: list[list[int]] = field(default_factory=list)
    length: list[float | None] = field(default_factory=list)
    support: list[float | None] = field(default_factory=list)
    tip_name: list[str | None] = field(default_factory=list)
    root: int = 0

    def add(self, name: str | None = None, length: float | None = None, support: float | None = None) -> int:
        self.children.append([])
        self.length.append(length)
        self.support.append(support)
        self.tip_name.append(name)
        return len(self.children) - 1

    def preorder(self) -> list[int]:
        out: list[int] = []
        stack = [self.root]
        while stack:
            x = stack.pop()
            out.append(x)
            stack.extend(reversed(self.children[x]))
        return out

    def tip_names(self) -> list[str]:
        return [self.tip_name[x] for x in self.preorder() if not self.children[x]]

    def to_phylonode(self) -> PhyloNode:
        made: dict[int, PhyloNode] = {}
        for x in reversed(self.preorder()):
            kids = [made.pop(c) for c in self.children[x]]
            name = self.tip_name[x] if not kids else None
            made[x] = PhyloNode(name, kids, self.length[x], self.support[x])
        root = made[self.root]
        counter = 0
        for node in root.preorder():
            if node.children:
                if node is root:
                    node.name = "root"
                else:
                    node.name = f"edge.{counter}"
                    counter += 1
        root.length = None
        return root

    def to_newick(self, with_lengths: bool = True, with_support: bool = False) -> str:
        text: dict[int, str] = {}
        for x in reversed(self.preorder()):
            if self.children[x]:
                s = "(" + ",".join(text.pop(c) for c in self.children[x]) + ")"
                if with_support and self.support[x] is not None and x != self.root:
                    s += repr(self.support[x])
            else:
                s = self.tip_name[x]
            if with_lengths and self.length[x] is not None and x != self.root:
                s += f":{self.length[x]!r}"
            text[x] = s
        return text[self.root] + ";"


# ---------------------------------------------------------------------------------------------
# model tree
# ---------------------------------------------------------------------------------------------
def birth_death_tree(n: int, rng: np.random.RandomState, birth: float = 1.0, death: float = 0.5) -> Topology:
    """Reconstructed birth-death tree with exactly ``n`` extant tips named t0..t{n-1}."""
    while True:
        parent = [-1]
        born = [0.0]
        died: list[float | None] = [None]  # time the lineage split or went extinct
        kids: list[list[int]] = [[]]
        alive = [0]
        now = 0.0
        while 0 < len(alive) < n:
            now += rng.exponential(1.0 / ((birth + death) * len(alive)))
            pick = rng.randint(len(alive))
            x = alive[pick]
            alive[pick] = alive[-1]
            alive.pop()
            died[x] = now
            if rng.random_sample() < birth / (birth + death):
                for _ in range(2):
                    parent.append(x)
                    born.append(now)
                    died.append(None)
                    kids.append([])
                    kids[x].append(len(parent) - 1)
                    alive.append(len(parent) - 1)
        if len(alive) == n:
            break
    now += rng.exponential(1.0 / ((birth + death) * n))
    for x in alive:
        died[x] = now
    # prune extinct lineages: keep nodes with an extant descendant, merge unary nodes
    extant = set(alive)
    has = [False] * len(parent)
    for x in range(len(parent) - 1, -1, -1):
        if x in extant or any(has[c] for c in kids[x]):
            has[x] = True
    topo = Topology()
    order = rng.permutation(n)
    next_tip = 0
    built: dict[int, tuple[int, float]] = {}  # sim node -> (topology node, accumulated length)
    for x in range(len(parent) - 1, -1, -1):
        if not has[x]:
            continue
        span = died[x] - born[x]
        live = [built.pop(c) for c in kids[x] if has[c]]
        if not live:
            node = topo.add(f"t{order[next_tip]}", None)
            next_tip += 1
            built[x] = (node, span)
        elif len(live) == 1:
            built[x] = (live[0][0], live[0][1] + span)
        else:
            node = topo.add(None, None)
            for child, acc in live:
                topo.length[child] = max(acc, 1e-6)
                topo.children[node].append(child)
            built[x] = (node, span)
    topo.root = built[0][0]
    return topo


# ---------------------------------------------------------------------------------------------
# inducing the model tree on a sample of tips, in O(k) after an O(n log n) index
# ---------------------------------------------------------------------------------------------
class _ModelIndex:
    def __init__(self, model: Topology) -> None:
        order = model.preorder()
        depth = {model.root: 0}
        cum = {model.root: 0.0}
        for x in order:
            for c in model.children[x]:
                depth[c] = depth[x] + 1
                cum[c] = cum[x] + (model.length[c] or 0.0)
        self.model = model
        self.tips = [x for x in order if not model.children[x]]
        self.tip_pos = {x: i for i, x in enumerate(self.tips)}
        self.depth = depth
        self.cum = cum
        # LCA of consecutive tips: the node the walk turns around at
        adj_node: list[int] = []
        stack = [[model.root, 0]]
        turn = model.root
        seen_tip = False
        while stack:
            frame = stack[-1]
            x, ci = frame
            if ci == len(model.children[x]):
                stack.pop()
                continue
            frame[1] = ci + 1
            if ci:
                turn = x
            c = model.children[x][ci]
            if model.children[c]:
                stack.append([c, 0])
            else:
                if seen_tip:
                    adj_node.append(turn)
                seen_tip = True
        self.adj_node = np.asarray(adj_node, dtype=np.int64)
        key = np.asarray([depth[x] for x in adj_node], dtype=np.int64) * (len(order) + 1) + self.adj_node
        levels = [key]
        span = 1
        while 2 * span <= len(key):
            prev = levels[-1]
            levels.append(np.minimum(prev[:-span], prev[span:]))
            span *= 2
        self.levels = levels
        self.stride = len(order) + 1
        # clade bookkeeping for the clade-based sampler
        self.lo: dict[int, int] = {}
        self.hi: dict[int, int] = {}
        for x in reversed(order):
            if model.children[x]:
                self.lo[x] = self.lo[model.children[x][0]]
                self.hi[x] = self.hi[model.children[x][-1]]
            else:
                self.lo[x] = self.tip_pos[x]
                self.hi[x] = self.tip_pos[x] + 1

    def lca_between(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        """LCA node of tips at positions left[i] < right[i] (range minimum over adj_node)."""
        width = right - left
        level = np.floor(np.log2(width)).astype(np.int64)
        out = np.empty(len(left), dtype=np.int64)
        for lv in np.unique(level):
            sel = level == lv
            table = self.levels[lv]
            a = table[left[sel]]
            b = table[right[sel] - (1 << lv)]
            out[sel] = np.minimum(a, b) % self.stride
        return out

    def induce(self, positions: np.ndarray, keep_lengths: bool) -> Topology:
        positions = np.sort(positions)
        topo = Topology()
        leaves = []
        for p in positions:
            x = self.tips[p]
            leaves.append((topo.add(self.model.tip_name[x], None), x))
        if len(positions) == 1:
            topo.root = leaves[0][0]
            return topo
        between = self.lca_between(positions[:-1], positions[1:])
        open_nodes: list[tuple[int, int]] = []  # (topology node, model node), depth increasing
        cur = leaves[0]
        for i, u in enumerate(between):
            u = int(u)
            d = self.depth[u]
            while open_nodes and self.depth[open_nodes[-1][1]] > d:
                top = open_nodes.pop()
                self._attach(topo, top, cur, keep_lengths)
                cur = top
            if open_nodes and open_nodes[-1][1] == u:
                self._attach(topo, open_nodes[-1], cur, keep_lengths)
            else:
                fresh = (topo.add(None, None), u)
                self._attach(topo, fresh, cur, keep_lengths)
                open_nodes.append(fresh)
            cur = leaves[i + 1]
        while open_nodes:
            top = open_nodes.pop()
            self._attach(topo, top, cur, keep_lengths)
            cur = top
        topo.root = cur[0]
        return topo

    def _attach(self, topo: Topology, parent, child, keep_lengths: bool) -> None:
        topo.children[parent[0]].append(child[0])
        if keep_lengths:
            topo.length[child[0]] = max(self.cum[child[1]] - self.cum[parent[1]], 1e-6)


def _nni(topo: Topology, moves: int, rng: np.random.RandomState) -> None:
    """Random nearest-neighbour interchanges on internal edges that do not touch the root."""
    parent = {}
    for x in topo.preorder():
        for c in topo.children[x]:
            parent[c] = x
    inner = [x for x in parent if topo.children[x] and parent[x] != topo.root]
    if not inner:
        return
    for _ in range(moves):
        v = inner[rng.randint(len(inner))]
        u = parent[v]
        siblings = [c for c in topo.children[u] if c != v]
        if not siblings:
            continue
        s = siblings[rng.randint(len(siblings))]
        c = topo.children[v][rng.randint(len(topo.children[v]))]
        iu = topo.children[u].index(s)
        iv = topo.children[v].index(c)
        topo.children[u][iu] = c
        topo.children[v][iv] = s
        parent[c] = u
        parent[s] = v
        # the candidate set changes only for the two moved subtrees' roots
        for y in (c, s):
            if topo.children[y]:
                ok = parent[y] != topo.root
                if ok and y not in inner:
                    inner.append(y)


@dataclass
class Problem:
    """One synthetic supertree problem."""

    n: int
    model: Topology
    sources: list[Topology]
    weights: list[float] | None
    weighting: str
    seed: int

    def phylonodes(self) -> list[PhyloNode]:
        return [s.to_phylonode() for s in self.sources]

    def names(self) -> list[str]:
        """Sorted names of every taxon that occurs in a source tree (global taxon id = index)."""
        found: set[str] = set()
        for s in self.sources:
            found.update(name for name, kids in zip(s.tip_name, s.children, strict=True) if not kids)
        return sorted(found)

    def forest_arrays(self) -> dict:
        """The source trees as the flat pre-order arrays of ``scs_forest_create`` (include/scs_b200.h),
        built straight from the child lists without per-node Python objects."""
        names = self.names()
        taxon_id = {name: i for i, name in enumerate(names)}
        keep_lengths = self.weighting == "branch"
        keep_support = self.weighting == "bootstrap"
        offsets = [0]
        parent: list[int] = []
        taxon: list[int] = []
        length: list[float] = []
        support: list[float] = []
        nan = float("nan")
        for s in self.sources:
            base = len(parent)
            stack = [(s.root, -1)]
            while stack:
                x, up = stack.pop()
                k = len(parent) - base
                parent.append(up)
                kids = s.children[x]
                taxon.append(-1 if kids else taxon_id[s.tip_name[x]])
                ln = s.length[x] if keep_lengths else None
                sp = s.support[x] if keep_support else None
                length.append(nan if ln is None else float(ln))
                support.append(nan if sp is None else float(sp))
                stack.extend((c, k) for c in reversed(kids))
            offsets.append(len(parent))
        weights = [1.0] * len(self.sources) if self.weights is None else list(self.weights)
        return {
            "names": names,
            "node_offsets": np.asarray(offsets, dtype=np.int64),
            "parent": np.asarray(parent, dtype=np.int32),
            "length": np.asarray(length, dtype=np.float64),
            "support": np.asarray(support, dtype=np.float64),
            "taxon": np.asarray(taxon, dtype=np.int32),
            "weights": np.asarray(weights, dtype=np.float64),
        }

    def newick_lines(self) -> list[str]:
        return [s.to_newick(with_lengths=self.weighting == "branch", with_support=self.weighting == "bootstrap")
                for s in self.sources]  # fmt: skip


def _source_tree(index: _ModelIndex, clades: list[int], t: int, n: int, weighting: str, nni_fraction: float,
                 rng: np.random.RandomState) -> Topology:
    """Source tree ``t`` of the recipe, drawing from ``rng``."""
    cap = math.ceil(0.25 * n)
    if t == 0:
        k = max(4, math.ceil(0.2 * n))
        positions = rng.choice(n, size=min(k, n), replace=False)
    else:
        x = clades[rng.randint(len(clades))]
        lo, hi = index.lo[x], index.hi[x]
        inside = np.arange(lo, hi)
        chosen = inside[rng.random_sample(len(inside)) < 0.5]
        if len(chosen) > cap:
            chosen = rng.choice(chosen, size=cap, replace=False)
        if len(chosen) < 4:
            chosen = rng.choice(inside, size=min(4, len(inside)), replace=False)
        outside = n - (hi - lo)
        if outside > 0:
            o = rng.randint(outside)
            o = o if o < lo else o + (hi - lo)
            chosen = np.append(chosen, o)
        positions = chosen
    topo = index.induce(np.asarray(positions, dtype=np.int64), weighting == "branch")
    tips = sum(1 for c in topo.children if not c)
    _nni(topo, math.ceil(nni_fraction * tips), rng)
    if weighting == "bootstrap":
        for x in range(len(topo.children)):
            if topo.children[x]:
                topo.support[x] = float(rng.randint(50, 101))
    return topo


def _clades(index: _ModelIndex, model: Topology, n: int) -> list[int]:
    min_clade = math.ceil(0.05 * n)
    clades = [x for x in index.lo if index.hi[x] - index.lo[x] >= min_clade and x != model.root]
    return clades or [model.root]


def make_problem(
    n: int,
    num_trees: int,
    weighting: str,
    seed: int,
    *,
    tree_weights: bool = False,
    nni_fraction: float = 0.05,
) -> Problem:
    """The SURVEY.md section 8d recipe; everything drawn from ``RandomState(seed)``."""
    rng = np.random.RandomState(seed)
    model = birth_death_tree(n, rng)
    index = _ModelIndex(model)
    clades = _clades(index, model, n)
    sources = [_source_tree(index, clades, t, n, weighting, nni_fraction, rng) for t in range(num_trees)]
    weights = None
    if tree_weights:
        weights = [float(w) for w in rng.uniform(0.5, 2.0, size=num_trees)]
    return Problem(n=n, model=model, sources=sources, weights=weights, weighting=weighting, seed=seed)


# ---- large workloads: per-tree random streams, trees generated and flattened by a pool of processes ----
_POOL_STATE: dict = {}


def _flat_chunk(trees: tuple[int, int]):
    """Worker: source trees [t0, t1) as flat pre-order arrays; tips carry the model tip number."""
    st = _POOL_STATE
    parent: list[int] = []
    tip: list[int] = []
    length: list[float] = []
    support: list[float] = []
    sizes: list[int] = []
    nan = float("nan")
    keep_lengths = st["weighting"] == "branch"
    keep_support = st["weighting"] == "bootstrap"
    for t in range(*trees):
        rng = np.random.RandomState([st["seed"], t + 1])
        s = _source_tree(st["index"], st["clades"], t, st["n"], st["weighting"], st["nni_fraction"], rng)
        base = len(parent)
        stack = [(s.root, -1)]
        while stack:
            x, up = stack.pop()
            k = len(parent) - base
            parent.append(up)
            kids = s.children[x]
            tip.append(-1 if kids else int(s.tip_name[x][1:]))
            ln = s.length[x] if keep_lengths else None
            sp = s.support[x] if keep_support else None
            length.append(nan if ln is None else float(ln))
            support.append(nan if sp is None else float(sp))
            stack.extend((c, k) for c in reversed(kids))
        sizes.append(len(parent) - base)
    return (np.asarray(sizes, dtype=np.int64), np.asarray(parent, dtype=np.int32), np.asarray(tip, dtype=np.int32),
            np.asarray(length, dtype=np.float64), np.asarray(support, dtype=np.float64))  # fmt: skip


def make_forest_arrays(
    n: int,
    num_trees: int,
    weighting: str,
    seed: int,
    *,
    tree_weights: bool = False,
    nni_fraction: float = 0.05,
    workers: int | None = None,
) -> dict:
    """Same recipe and output as ``make_problem(...).forest_arrays()``, for workloads too large to generate
    on one core: the model tree comes from ``RandomState(seed)``, source tree ``t`` from its own stream
    ``RandomState([seed, t + 1])``, so the trees can be made by a pool of forked processes and the result does
    not depend on the number of workers.  (A different problem instance than ``make_problem`` with the same
    seed, which draws everything from one stream.)"""
    import multiprocessing as mp
    import os

    rng = np.random.RandomState(seed)
    model = birth_death_tree(n, rng)
    index = _ModelIndex(model)
    _POOL_STATE.update(index=index, clades=_clades(index, model, n), n=n, weighting=weighting, seed=seed,
                       nni_fraction=nni_fraction)  # fmt: skip
    workers = max(1, workers or min(32, os.cpu_count() or 1))
    step = max(1, min(64, num_trees // (4 * workers) or 1))
    chunks = [(t0, min(num_trees, t0 + step)) for t0 in range(0, num_trees, step)]
    if workers == 1:
        parts = [_flat_chunk(c) for c in chunks]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_flat_chunk, chunks, chunksize=1)
    _POOL_STATE.clear()
    sizes = np.concatenate([p[0] for p in parts])
    tip = np.concatenate([p[2] for p in parts])
    found = np.unique(tip[tip >= 0])
    names = sorted(f"t{i}" for i in found)
    rank = np.full(n, -1, dtype=np.int32)
    for i, name in enumerate(names):
        rank[int(name[1:])] = i
    taxon = np.where(tip >= 0, rank[np.maximum(tip, 0)], -1).astype(np.int32)
    weights = rng.uniform(0.5, 2.0, size=num_trees) if tree_weights else np.ones(num_trees)
    return {
        "names": names,
        "node_offsets": np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
        "parent": np.concatenate([p[1] for p in parts]),
        "length": np.concatenate([p[3] for p in parts]),
        "support": np.concatenate([p[4] for p in parts]),
        "taxon": taxon,
        "weights": np.asarray(weights, dtype=np.float64),
    }
