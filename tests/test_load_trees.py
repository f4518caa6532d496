"""CPU: ``load_trees`` (ref: src/sc_supertree/load.py:7-23) returns the reference's list of tree objects, parsed
natively: each element builds its node objects when first looked at (``load.LoadedTree``), and
``construct_supertree`` takes the natively parsed forest when nobody has."""

from __future__ import annotations

import copy
import json
import pickle

import numpy as np
import pytest

from helpers import GOLDEN, load_ctrace
from spectralclustersupertree_b200 import construct_supertree, load_trees
from spectralclustersupertree_b200.engine import Forest
from spectralclustersupertree_b200.load import LoadedTree, untouched_forest
from spectralclustersupertree_b200.tree import NewickError, PhyloNode, make_tree


def write(tmp_path, lines, name="source.tre"):
    path = tmp_path / name
    path.write_text("\n".join(lines) + "\n")
    return path


def same_forest(a: Forest, b: Forest) -> None:
    assert a.names == b.names and a.num_trees == b.num_trees
    assert np.array_equal(a.weights(), b.weights())
    for t in range(a.num_trees):
        for x, y in zip(a.tree_arrays(t), b.tree_arrays(t), strict=True):
            assert np.array_equal(x, y, equal_nan=True), t


def same_tree(a: PhyloNode, b: PhyloNode) -> None:
    nodes_a, nodes_b = list(a.preorder()), list(b.preorder())
    assert len(nodes_a) == len(nodes_b)
    for x, y in zip(nodes_a, nodes_b, strict=True):
        assert (x.name, x.length, x.support, len(x.children)) == (y.name, y.length, y.support, len(y.children))
        assert (x.parent is None) == (y.parent is None)


@pytest.mark.parametrize("source", ["fixture_supertriplets", "fixture_dcm_iq", "untidy_branch", "untidy_bootstrap"])
def test_loaded_list_is_the_natively_parsed_forest_until_someone_looks(tmp_path, source):
    if source.startswith("fixture_"):
        lines = json.loads((GOLDEN / f"{source}.json").read_text())["trees"]
    else:
        lines = load_ctrace(source)["lines"]
    trees = load_trees(write(tmp_path, lines))
    assert type(trees) is list and len(trees) == len(lines)
    assert all(type(t) is LoadedTree and isinstance(t, PhyloNode) for t in trees)
    eager = [make_tree(s) for s in lines]
    same_forest(untouched_forest(trees, None), Forest.from_trees(eager, [1.0] * len(eager)))
    weights = [0.5 + 0.25 * i for i in range(len(lines))]
    same_forest(untouched_forest(trees, weights), Forest.from_trees(eager, weights))
    # a different list of the same objects is not "what load_trees returned"
    assert untouched_forest(trees[:-1], None) is None
    assert untouched_forest(trees[::-1], None) is None
    assert untouched_forest([*trees[:-1], eager[-1]], None) is None
    assert untouched_forest(tuple(trees), None) is not None
    # looking at one tree builds that tree, exactly as make_tree does, and ends the shortcut
    same_tree(trees[1], eager[1])
    assert all(child.parent is trees[1] for child in trees[1].children)
    assert untouched_forest(trees, None) is None
    for got, want in zip(trees, eager, strict=True):
        same_tree(got, want)
        assert got.get_newick(with_distances=True) == want.get_newick(with_distances=True)
    # and the generic route flattens the built objects to the same forest (changes made to them included)
    same_forest(Forest.from_trees(trees, weights), Forest.from_trees(eager, weights))
    tip = next(trees[0].iter_tips())
    tip.name = "renamed_tip"
    assert "renamed_tip" in Forest.from_trees(trees, weights).names


def test_assignment_copies_and_pickles_build_the_tree_first(tmp_path):
    lines = ["((a:1,b:2)90:3,(c:1,d:4):2);", "(a,(b,(c,d)));", "((a,b),(c,e));"]
    trees = load_trees(write(tmp_path, lines))
    trees[0].name = "mine"  # assignment before any read: the tree is built first, then the assignment applies
    assert trees[0].name == "mine" and len(trees[0].children) == 2 and trees[0].children[0].support == 90.0
    assert untouched_forest(trees, None) is None
    trees = load_trees(write(tmp_path, lines))
    for clone in (copy.copy(trees[1]), copy.deepcopy(trees[1]), pickle.loads(pickle.dumps(trees[1])), trees[1].copy()):
        assert sorted(clone.get_tip_names()) == ["a", "b", "c", "d"]
        assert clone.same_shape(make_tree(lines[1]))
    assert repr(trees[2]).startswith("Tree(")
    assert [t.is_tip() for t in trees] == [False, False, False]


def test_files_the_flat_store_refuses_and_syntax_errors(tmp_path):
    # the same taxon on two tips of one tree: the reference's loader accepts the file, so node objects right away
    trees = load_trees(write(tmp_path, ["((a,b),(c,d));", "(a,a,(b,c));"]))
    assert [type(t) for t in trees] == [PhyloNode, PhyloNode]
    assert trees[1].get_tip_names() == ["a", "a", "b", "c"]
    with pytest.raises(NewickError):
        load_trees(write(tmp_path, ["((a,b),(c,d));", "(a,b;"]))
    with pytest.raises(NewickError):  # an empty line is a syntax error, as in the reference (load.py:21-22)
        load_trees(write(tmp_path, ["(a,b);", "", "(c,d);"]))
    empty = tmp_path / "empty.tre"
    empty.write_text("")
    assert load_trees(empty) == []
    blanks = load_trees(write(tmp_path, ["  (a , b) ;  ", "\t(b,c);"]))  # lines are stripped before parsing
    assert [sorted(t.get_tip_names()) for t in blanks] == [["a", "b"], ["b", "c"]]


def test_construct_supertree_on_an_untouched_list_needs_no_node_objects(tmp_path):
    """Decisions taken on the host, before any device call: two taxa give a star (ref: scs.py:105-106) straight from
    the parsed names; a single tree takes the single-tree shortcut (ref: scs.py:96-98) through its node objects."""
    trees = load_trees(write(tmp_path, ["(a,b);", "(b,a);", "a;"]))
    assert construct_supertree(trees).sorted().get_newick() == "(a,b);"
    assert untouched_forest(trees, None) is not None  # still untouched
    one = load_trees(write(tmp_path, ["((a,b)x,(c,d)y);"]))
    assert construct_supertree(one).sorted().same_shape(make_tree("((a,b),(c,d));").sorted())
    with pytest.raises(ValueError, match=r"The number of trees \(3\) and tree weights \(1\) must match."):
        construct_supertree(trees, weights=[1.0])
    with pytest.raises(ValueError, match="Invalid weighting strategy selected: 'bogus'"):
        construct_supertree(trees, pcg_weighting="bogus")


def test_cogent3_trees_in_cogent3_tree_out(monkeypatch):
    """With real cogent3 trees in, the supertree is handed back as a cogent3 tree (``cogent3.make_tree`` of its
    Newick text); the package's own class otherwise, and when ``cogent3`` is only the oracle's import shim."""
    import sys
    import types

    from spectralclustersupertree_b200 import scs

    class Foreign(PhyloNode):  # stands for cogent3.core.tree.PhyloNode: same surface, another module
        __slots__ = ()

    Foreign.__module__ = "cogent3.core.tree"
    made = []

    def fake_make_tree(text):
        made.append(text)
        return ("cogent3 tree of", text)

    fake = types.ModuleType("cogent3")
    fake.make_tree = fake_make_tree
    monkeypatch.setitem(sys.modules, "cogent3", fake)
    ours = [make_tree("(a,b);"), make_tree("(b,a);")]
    theirs = [Foreign("root", [Foreign("a"), Foreign("b")]), Foreign("root", [Foreign("b"), Foreign("a")])]
    assert isinstance(construct_supertree(ours), PhyloNode) and not made
    assert construct_supertree(theirs) == ("cogent3 tree of", "(a,b);")
    assert construct_supertree(theirs[:1])[0] == "cogent3 tree of"  # the single-tree shortcut too
    fake._scs_b200_shim = True
    assert isinstance(construct_supertree(theirs), PhyloNode)
    assert scs._in_the_callers_class(ours[0], ours) is ours[0]
