"""CPU: the flat source-tree store (csrc/forest.cpp) against the PhyloNode restatement.

``Forest.induce`` must reproduce ``PhyloNode.get_sub_tree(names, ignore_missing=True,
as_rooted=True)`` as the reference uses it (ref: scs.py:444-453) -- including the order of the
floating-point additions when unary nodes are merged -- and ``Forest.tours`` must reproduce the
Python flattening of ``flatten.py`` for every weighting."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import load_case, parse
from spectralclustersupertree_b200.engine import Forest
from spectralclustersupertree_b200.flatten import flatten_trees
from spectralclustersupertree_b200.tree import make_tree


def assert_same_tours(a, b):
    assert a.n == b.n
    for field in ("leaf_offsets", "leaf_taxon", "adj_depth", "root_depth", "tree_weight"):
        assert np.array_equal(getattr(a, field), getattr(b, field)), field
    assert np.array_equal(a.adj_val, b.adj_val), "adj_val"  # bit-exact


@pytest.mark.parametrize(
    ("name", "weighting"),
    [("supertriplets", "depth"), ("supertriplets", "one"), ("dcm_iq", "branch"), ("s_200x40_bootstrap", "bootstrap"),
     ("s_300x40_branch_weighted", "branch"), ("c1_100x30_depth", "depth")],
)  # fmt: skip
def test_tours_match_python_flatten(name, weighting):
    case = load_case(name)
    trees = parse(case["lines"])
    tid = {x: i for i, x in enumerate(case["names"])}
    forest = Forest.from_trees(trees, case["weights"], case["names"])
    assert forest.num_trees == len(trees)
    assert np.array_equal(forest.taxa(), np.arange(len(case["names"])))
    assert_same_tours(forest.tours(weighting), flatten_trees(trees, case["weights"], weighting, tid))


@pytest.mark.parametrize("name", ["dcm_iq", "s_300x40_branch_weighted", "s_200x40_bootstrap", "supertriplets"])
def test_induce_matches_get_sub_tree(name):
    case = load_case(name)
    trees = parse(case["lines"])
    names = case["names"]
    forest = Forest.from_trees(trees, case["weights"], names)
    rng = np.random.RandomState(5)
    for frac in (0.7, 0.3, 0.05):
        keep = np.flatnonzero(rng.random_sample(len(names)) < frac).astype(np.int32)
        if len(keep) < 3:
            continue
        kept_names = {names[i] for i in keep}
        sub = forest.induce(keep)
        ref_trees, ref_weights = [], []
        for tree, w in zip(trees, case["weights"], strict=True):  # ref: scs.py:444-453
            if len(kept_names.intersection(tree.get_tip_names())) < 2:
                continue
            s = tree.get_sub_tree(kept_names, ignore_missing=True, as_rooted=True)
            s.name = "root"
            ref_trees.append(s)
            ref_weights.append(w)
        assert sub.num_trees == len(ref_trees)
        assert np.array_equal(sub.weights(), np.asarray(ref_weights))
        present = sub.taxa()
        local_names = [names[i] for i in present]
        tid = {x: i for i, x in enumerate(local_names)}
        for weighting in ("one", "depth", "branch", "bootstrap"):
            if weighting == "bootstrap" and case["weighting"] != "bootstrap":
                continue
            assert_same_tours(sub.tours(weighting), flatten_trees(ref_trees, ref_weights, weighting, tid))
        # a second restriction of the restricted forest behaves the same (the recursion nests them)
        keep2 = present[:: 2]
        if len(keep2) >= 3:
            sub2 = sub.induce(keep2)
            names2 = {names[i] for i in keep2}
            ref2 = [t.get_sub_tree(names2, ignore_missing=True, as_rooted=True) for t in ref_trees
                    if len(names2.intersection(t.get_tip_names())) >= 2]  # fmt: skip
            w2 = [w for t, w in zip(ref_trees, ref_weights, strict=True)
                  if len(names2.intersection(t.get_tip_names())) >= 2]  # fmt: skip
            tid2 = {names[x]: i for i, x in enumerate(sub2.taxa())}
            assert_same_tours(sub2.tours(case["weighting"]), flatten_trees(ref2, w2, case["weighting"], tid2))


def test_unary_nodes_polytomies_and_lone_tips():
    lines = ["((a:1,b:2):3,((c:1):2,(d:1,e:2,f:3):4):5);", "(a,(b,(c)));", "x;", "((x,a),b);"]
    trees = parse(lines)
    names = sorted({n for t in trees for n in t.get_tip_names()})
    tid = {x: i for i, x in enumerate(names)}
    forest = Forest.from_trees(trees, [1.0, 2.0, 3.0, 4.0], names)
    assert forest.num_trees == 4
    for weighting in ("one", "depth", "branch"):
        assert_same_tours(forest.tours(weighting), flatten_trees(trees, [1.0, 2.0, 3.0, 4.0], weighting, tid))
    sub = forest.induce(np.array([tid[x] for x in "acdx"], dtype=np.int32))
    ref = [t.get_sub_tree(set("acdx"), ignore_missing=True, as_rooted=True) for t in trees
           if len(set("acdx") & set(t.get_tip_names())) >= 2]  # fmt: skip
    assert sub.num_trees == len(ref) == 3
    tid2 = {names[x]: i for i, x in enumerate(sub.taxa())}
    assert_same_tours(sub.tours("branch"), flatten_trees(ref, [1.0, 2.0, 4.0], "branch", tid2))


def test_bootstrap_without_support_raises_type_error():
    trees = [make_tree("(a,(b,(c,d)));"), make_tree("(a,(b,c));")]
    names = ["a", "b", "c", "d"]
    forest = Forest.from_trees(trees, [1.0, 1.0], names)
    with pytest.raises(TypeError):
        forest.tours("bootstrap")


def test_forest_rejects_malformed_arrays():
    from spectralclustersupertree_b200.engine import ScsError

    with pytest.raises(ScsError):  # parent index not smaller than the node's own index
        Forest.from_arrays([0, 3], [-1, 2, 0], None, None, [-1, 0, 1], [1.0], ["a", "b"])
    with pytest.raises(ScsError):  # tip without a taxon id
        Forest.from_arrays([0, 3], [-1, 0, 0], None, None, [-1, 0, -1], [1.0], ["a", "b"])
    bad = {
        "internal node with a taxon id": ([0, 3], [-1, 0, 0], [1, 0, 1]),
        "taxon id out of range": ([0, 3], [-1, 0, 0], [-1, 0, 2]),
        "negative taxon id on a tip": ([0, 3], [-1, 0, 0], [-1, 0, -7]),
        "first node is not a root": ([0, 3], [0, 0, 0], [-1, 0, 1]),
        "negative parent below the root": ([0, 3], [-1, -1, 0], [-1, 0, 1]),
        "a tree without nodes": ([0, 0, 3], [-1, 0, 0], [-1, 0, 1]),
        "the last tree is wrong": ([0, 3, 6], [-1, 0, 0, -1, 0, 5], [-1, 0, 1, -1, 0, 1]),
    }
    for what, (offsets, parent, taxon) in bad.items():
        weights = [1.0] * (len(offsets) - 1)
        for copy in (False, True):
            with pytest.raises(ScsError, match="not a valid pre-order tree"):
                Forest.from_arrays(offsets, parent, None, None, taxon, weights, ["a", "b"], copy=copy)
            assert what
    # and the shapes that are fine: a lone tip, a unary chain, two trees
    ok = Forest.from_arrays([0, 1, 4, 7], [-1, -1, 0, 1, -1, 0, 0], None, None, [0, -1, -1, 1, -1, 0, 1], [1.0, 1.0, 1.0], ["a", "b"])
    assert ok.num_trees == 3 and ok.num_leaves == 3  # the lone tip is in no tour: 1 + 2 leaves


@pytest.mark.parametrize("name", ["dcm_iq", "s_300x40_branch_weighted", "c2_500x50_branch"])
@pytest.mark.parametrize("parts", [2, 5])
def test_induce_parts_equals_one_restriction_per_part(name, parts):
    """The batched restriction the recursion driver uses (all children of a node in one pass over the
    host threads) against ``induce`` part by part: same trees, same arrays, bit for bit."""
    case = load_case(name)
    names = case["names"]
    forest = Forest.from_trees(parse(case["lines"]), case["weights"], names)
    rng = np.random.RandomState(11)
    part = rng.randint(0, parts + 1, size=len(names)).astype(np.int32)
    part[part == parts] = -1  # some taxa belong to no part (stars / dropped)
    subs, present = forest.induce_parts(part, parts)
    assert len(subs) == parts
    seen = []
    for c, sub in enumerate(subs):
        want = forest.induce(np.flatnonzero(part == c).astype(np.int32))
        assert sub.num_trees == want.num_trees
        assert sub.num_leaves == want.num_leaves
        assert np.array_equal(sub.weights(), want.weights())
        for t in range(sub.num_trees):
            for got_arr, want_arr in zip(sub.tree_arrays(t), want.tree_arrays(t), strict=True):
                assert np.array_equal(got_arr, want_arr, equal_nan=True)
        assert_same_tours(sub.tours(case["weighting"]), want.tours(case["weighting"]))
        seen.append(want.taxa())
    assert np.array_equal(present, np.sort(np.concatenate(seen)))


def test_induce_parts_keeps_whole_trees_verbatim():
    """Parts that swallow whole source trees (the common case deep in the recursion): the batched restriction
    copies such a tree as it is, unless it has unary nodes to merge; either way the result equals ``induce``.
    Applied twice, so that restricted forests are restricted again."""
    lines = [
        "((a:1,b:2):0.5,(c:3,d:4)90:0.25,e:5);",  # whole in part 0
        "((a:1,(b:2):7):0.5,((c:3)));",  # unary nodes, whole in part 0: must still be merged
        "((f:1,g:2):1,(h:3,(i:4,j:5):6):2);",  # whole in part 1
        "((a:1,f:2):1,(b:3,g:4):2,c:1);",  # split between the parts
        "(a:1,b:1);",  # two tips, whole in part 0
        "(a:1,f:1);",  # one tip per part: dropped from both
        "k;",  # a lone tip of no part
    ]
    trees = parse(lines)
    names = sorted({x for t in trees for x in t.get_tip_names()})
    forest = Forest.from_trees(trees, [1.0, 2.0, 0.5, 1.5, 1.0, 1.0, 1.0], names)
    part = np.array([0 if x in "abcde" else (1 if x in "fghij" else -1) for x in names], dtype=np.int32)

    def check(src: Forest, part: np.ndarray, parts: int) -> list[Forest]:
        subs, present = src.induce_parts(part, parts)
        seen = []
        for c, sub in enumerate(subs):
            want = src.induce(np.flatnonzero(part == c).astype(np.int32))
            assert sub.num_trees == want.num_trees
            assert np.array_equal(sub.weights(), want.weights())
            for t in range(sub.num_trees):
                for got_arr, want_arr in zip(sub.tree_arrays(t), want.tree_arrays(t), strict=True):
                    assert np.array_equal(got_arr, want_arr, equal_nan=True), (c, t)
            assert_same_tours(sub.tours("branch"), want.tours("branch"))
            seen.append(want.taxa())
        assert np.array_equal(present, np.sort(np.concatenate(seen)))
        return subs

    first = check(forest, part, 2)
    assert first[0].num_trees == 4 and first[1].num_trees == 2
    # again: part 0's forest split into {a, b} | {c, d, e}; its whole-tree copies are restricted for real now
    again = np.array([0 if x in "ab" else (1 if x in "cde" else -1) for x in names], dtype=np.int32)
    check(first[0], again, 2)


def test_a_taxon_on_two_tips_of_one_tree_is_rejected():
    """The graph kernels give every leaf of a tree its own column (no atomics on W): a source tree that
    carries the same taxon twice is refused when the forest is made, with a message saying so."""
    from spectralclustersupertree_b200.engine import ScsError

    trees = [make_tree("((a,b),(c,d));"), make_tree("(a,a,(b,c));")]
    with pytest.raises(ScsError, match="more than one tip"):
        Forest.from_trees(trees, [1.0, 1.0])
    with pytest.raises(Exception, match="more than one tip"):
        Forest.from_newick("((a,b),(c,d));\n(a,a,(b,c));\n")


def test_forest_view_equals_forest_copy():
    """``scs_forest_create_view`` (the forest refers to the caller's per-node arrays) against ``scs_forest_create``
    (the library's own copy): same tours, same restrictions, same validation errors; a forest created without
    lengths / supports still gets its own NaN arrays."""
    from spectralclustersupertree_b200.engine import ScsError
    from spectralclustersupertree_b200.synthetic import make_problem

    arrays = make_problem(400, 60, "branch", 11, tree_weights=True).forest_arrays()

    def forest(copy, **override):
        a = {**arrays, **override}
        return Forest.from_arrays(a["node_offsets"], a["parent"], a["length"], a["support"], a["taxon"], a["weights"],
                                  a["names"], copy=copy)  # fmt: skip

    owned, view = forest(True), forest(False)
    for weighting in ("one", "depth", "branch"):
        assert_same_tours(owned.tours(weighting), view.tours(weighting))
    keep = np.arange(0, len(arrays["names"]), 3, dtype=np.int32)
    a, b = owned.induce(keep), view.induce(keep)
    assert a.num_trees == b.num_trees
    assert_same_tours(a.tours("branch"), b.tours("branch"))
    # the arrays the caller does not have are made by the library in either mode
    bare = forest(False, length=None, support=None)
    assert_same_tours(bare.tours("depth"), owned.tours("depth"))
    assert np.all(bare.tours("branch").adj_val[bare.tours("branch").adj_depth > 0] >= 1.0)  # missing length counts 1
    # validation sees the same input whoever owns it
    broken = arrays["parent"].copy()
    broken[5] = 7  # a parent index must be smaller than the node's own
    for copy in (True, False):
        with pytest.raises(ScsError):
            forest(copy, parent=broken)
    twice = arrays["taxon"].copy()
    tips = np.flatnonzero(twice[: arrays["node_offsets"][1]] >= 0)
    twice[tips[1]] = twice[tips[0]]  # the same taxon on two tips of the first tree
    for copy in (True, False):
        with pytest.raises(ScsError, match="more than one tip"):
            forest(copy, taxon=twice)


def test_from_trees_without_names_sorts_the_tip_names_it_finds():
    """``construct_supertree`` lets the flattener collect the taxon names (no ``get_tip_names`` pass of its own,
    ref: scs.py:100-103): taxon id = rank of the name in the sorted list, whatever order the tips are met in."""
    case = load_case("s_300x40_branch_weighted")
    trees = parse(case["lines"])
    given = Forest.from_trees(trees, case["weights"], case["names"])
    found = Forest.from_trees(trees, case["weights"])
    assert found.names == sorted({x for t in trees for x in t.get_tip_names()}) == list(case["names"])
    for t in range(given.num_trees):
        for a, b in zip(given.tree_arrays(t), found.tree_arrays(t), strict=True):
            assert np.array_equal(a, b, equal_nan=True)


def test_two_taxa_give_a_star_without_touching_the_gpu():
    """ref: scs.py:105-106 (and :96-98 for a single tree): decided on the host, before any device call."""
    from spectralclustersupertree_b200 import construct_supertree

    assert construct_supertree([make_tree("(a,b);"), make_tree("(b,a);")]).sorted().get_newick() == "(a,b);"
    assert construct_supertree([make_tree("(b,a);"), make_tree("b;")], weights=[1, 2]).sorted().get_newick() == "(a,b);"
    single = construct_supertree([make_tree("((a,b)x,(c,d)y);")])
    assert single.sorted().same_shape(make_tree("((a,b),(c,d));").sorted())


def test_random_trees_restrict_like_get_sub_tree():
    """Property test: random trees with unary chains, polytomies, missing lengths and supports.  ``induce`` and
    ``induce_parts`` must give, node for node and bit for bit (lengths are sums whose operand order matters), the
    arrays of ``get_sub_tree(names, ignore_missing=True, as_rooted=True)`` applied to every tree (ref: scs.py:444-453),
    and restricting the restricted forest again must still agree."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    from spectralclustersupertree_b200.tree import PhyloNode

    taxa = [f"t{i}" for i in range(12)]
    lengths = st.one_of(st.none(), st.floats(min_value=1e-3, max_value=10.0, allow_nan=False))
    supports = st.one_of(st.none(), st.integers(min_value=0, max_value=100).map(float))

    @st.composite
    def tree(draw):
        tips = draw(st.lists(st.sampled_from(taxa), min_size=1, max_size=10, unique=True))
        nodes = [PhyloNode(name, None, draw(lengths), None) for name in tips]
        while len(nodes) > 1:
            # join 1..3 of the current subtrees under a new internal node (1 => a unary node)
            k = draw(st.integers(min_value=1, max_value=min(3, len(nodes))))
            picked = [nodes.pop(draw(st.integers(min_value=0, max_value=len(nodes) - 1))) for _ in range(k)]
            if k == 1 and not nodes and draw(st.booleans()):
                return picked[0]
            nodes.append(PhyloNode("", picked, draw(lengths), draw(supports)))
            if k == 1 and len(nodes) == 1 and draw(st.booleans()):
                break  # a unary root
        return nodes[0]

    def arrays_of(forest):
        return [forest.tree_arrays(t) for t in range(forest.num_trees)]

    def reference_restriction(trees, weights, kept_names):
        out, out_w = [], []
        for t, w in zip(trees, weights, strict=True):
            if len(kept_names.intersection(t.get_tip_names())) < 2:
                continue
            s = t.get_sub_tree(kept_names, ignore_missing=True, as_rooted=True)
            s.name = "root"
            out.append(s)
            out_w.append(w)
        return out, out_w

    def same(got: Forest, want_trees, want_weights):
        assert got.num_trees == len(want_trees)
        if not want_trees:
            return
        assert np.array_equal(got.weights(), np.asarray(want_weights, dtype=np.float64))
        want = Forest.from_trees(want_trees, want_weights, taxa)
        for g, w in zip(arrays_of(got), arrays_of(want), strict=True):
            gp, gl, gs, gt = g
            wp, wl, ws, wt = w
            assert np.array_equal(gp, wp) and np.array_equal(gt, wt)
            # the new root's own length is dropped by both; every other node bit for bit
            assert np.array_equal(gl[1:], wl[1:], equal_nan=True)
            internal = gt < 0
            assert np.array_equal(gs[internal][1:], ws[internal][1:], equal_nan=True)

    @settings(max_examples=150, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(tree(), min_size=1, max_size=5), st.lists(st.integers(0, 2), min_size=12, max_size=12))
    def run(trees, assignment):
        weights = [1.0 + 0.5 * i for i in range(len(trees))]
        forest = Forest.from_trees(trees, weights, taxa)
        part = np.asarray(assignment, dtype=np.int32)
        part[part == 2] = -1
        subs, _ = forest.induce_parts(part, 2)
        for c in (0, 1):
            kept = {taxa[i] for i in np.flatnonzero(part == c)}
            want_trees, want_weights = reference_restriction(trees, weights, kept)
            same(subs[c], want_trees, want_weights)
            same(forest.induce(np.flatnonzero(part == c).astype(np.int32)), want_trees, want_weights)
            # once more on the restricted forest: every other kept taxon
            again = sorted(kept)[::2]
            if len(again) >= 2 and want_trees:
                ids = np.asarray([taxa.index(x) for x in again], dtype=np.int32)
                t2, w2 = reference_restriction(want_trees, want_weights, set(again))
                same(subs[c].induce(ids), t2, w2)

    run()


@pytest.mark.parametrize("case", ["branch", "bootstrap", "depth", "caterpillar"])
def test_untidy_trees_flatten_and_parse_alike(case):
    """The untidy golden cases (unary chains, polytomies, missing lengths, unary roots, two-tip trees, lone tips):
    the host forest's tours equal the Python flattening, and the native Newick parser builds the forest that
    ``make_tree`` + ``Forest.from_trees`` build."""
    from helpers import load_ctrace

    ctrace = load_ctrace(f"untidy_{case}")
    trees = parse(ctrace["lines"])
    names = sorted({x for t in trees for x in t.get_tip_names()})
    tid = {x: i for i, x in enumerate(names)}
    forest = Forest.from_trees(trees, ctrace["weights"], names)
    assert_same_tours(forest.tours(ctrace["weighting"]), flatten_trees(trees, ctrace["weights"], ctrace["weighting"], tid))
    parsed = Forest.from_newick("\n".join(ctrace["lines"]) + "\n")
    assert parsed.names == names and parsed.num_trees == forest.num_trees
    for t in range(forest.num_trees):
        for a, b in zip(parsed.tree_arrays(t), forest.tree_arrays(t), strict=True):
            assert np.array_equal(a, b, equal_nan=True), t


def test_flattener_reads_slots_directly_and_other_classes_generically():
    """csrc/fastflatten.c reads ``children`` / ``name`` / ``length`` / ``support`` straight out of the slots of a node
    class that declares them in ``__slots__`` (the package's PhyloNode) and goes through ``getattr`` for anything else
    (subclasses, classes with a ``__dict__``, tuples of children, cogent3's own class): same arrays either way, same
    errors for an unset name and a length that is not a number."""
    from spectralclustersupertree_b200.tree import PhyloNode

    class Sub(PhyloNode):
        pass

    class Plain:
        def __init__(self, name, kids=(), length=None):
            self.name, self.children, self.length = name, list(kids), length

        def __iter__(self):
            return iter(self.children)

    class TupleKids:
        __slots__ = ("children", "length", "name", "support")

        def __init__(self, name, kids=(), length=None):
            self.name, self.children, self.length, self.support = name, tuple(kids), length, None

    def shaped(cls):  # ((a:1,b:2):3,(c:1,d:4):2) in class `cls`
        def node(name, kids=(), length=None):
            made = cls(name, list(kids)) if cls in (PhyloNode, Sub) else cls(name, kids)
            made.length = length
            return made

        return node("root", [node("", [node("a", (), 1.0), node("b", (), 2.0)], 3.0),
                             node("", [node("c", (), 1.0), node("d", (), 4.0)], 2.0)])  # fmt: skip

    want = Forest.from_trees([make_tree("((a:1,b:2):3,(c:1,d:4):2);")], [1.0])
    for cls in (PhyloNode, Sub, Plain, TupleKids):
        got = Forest.from_trees([shaped(cls)], [1.0])
        assert got.names == want.names
        for a, b in zip(got.tree_arrays(0), want.tree_arrays(0), strict=True):
            assert np.array_equal(a, b, equal_nan=True), cls.__name__
    # classes mixed within one call: the first node's class gets the direct path, the others the generic one
    mixed = Forest.from_trees([shaped(PhyloNode), shaped(Plain), shaped(Sub)], [1.0, 1.0, 1.0])
    for t in range(3):
        for a, b in zip(mixed.tree_arrays(t), want.tree_arrays(0), strict=True):
            assert np.array_equal(a, b, equal_nan=True)

    class Bare:
        __slots__ = ("children", "length", "name", "support")

    bare = Bare()
    bare.children = []
    with pytest.raises(AttributeError):
        Forest.from_trees([bare], [1.0])
    odd = PhyloNode("a")
    odd.length = "long"
    with pytest.raises(TypeError):
        Forest.from_trees([PhyloNode("r", [odd, PhyloNode("b")])], [1.0])


def test_validation_agrees_with_a_plain_python_checker_on_corrupted_forests():
    """``scs_forest_create`` accumulates its checks as flags (no early exit): whatever single entry of a valid forest
    is overwritten, it must accept or refuse exactly as the rule it implements, spelled out here in plain Python."""
    from spectralclustersupertree_b200.engine import ScsError

    def valid(offsets, parent, taxon, num_taxa) -> bool:
        for t in range(len(offsets) - 1):
            base, count = offsets[t], offsets[t + 1] - offsets[t]
            if count < 1 or parent[base] != -1:
                return False
            seen = set()
            for k in range(count):
                p = parent[base + k]
                if k >= 1 and not 0 <= p < k:
                    return False
                tip = k + 1 >= count or parent[base + k + 1] != k
                x = taxon[base + k]
                if tip:
                    if not 0 <= x < num_taxa or x in seen:
                        return False
                    seen.add(x)
                elif x != -1:
                    return False
        return True

    case = load_case("supertriplets")
    names = case["names"]
    good = Forest.from_trees(parse(case["lines"]), case["weights"], names)
    arrays = [good.tree_arrays(t) for t in range(good.num_trees)]
    offsets = np.concatenate([[0], np.cumsum([len(a[0]) for a in arrays])]).astype(np.int64)
    parent = np.concatenate([a[0] for a in arrays]).astype(np.int32)
    taxon = np.concatenate([a[3] for a in arrays]).astype(np.int32)
    weights = np.ones(good.num_trees)
    assert valid(offsets, parent, taxon, len(names))
    rng = np.random.RandomState(17)
    refused = accepted = 0
    for _ in range(400):
        p, x = parent.copy(), taxon.copy()
        at = int(rng.randint(len(p)))
        if rng.random_sample() < 0.5:
            p[at] = int(rng.randint(-2, 12)) if rng.random_sample() < 0.7 else int(rng.randint(-2, len(p)))
        else:
            x[at] = int(rng.choice([-1, -5, 0, int(rng.randint(len(names))), len(names), len(names) + 3]))
        want = valid(offsets, p, x, len(names))
        for copy in (False, True):
            try:
                Forest.from_arrays(offsets, p, None, None, x, weights, names, copy=copy).close()
                got = True
            except ScsError:
                got = False
            assert got == want, (at, int(p[at]), int(x[at]), want)
        refused += not want
        accepted += want
    assert refused > 100 and accepted > 20
