"""GPU: the cogent3-app surface (ref: src/sc_supertree/_app.py:34-96, tests/test_app.py) on the stand-in apps of
``spectralclustersupertree_b200._app`` (cogent3 itself is not installed here; with it the same names are real
``define_app`` apps)."""

from __future__ import annotations

import json

import pytest

from helpers import GOLDEN, kat_cases, rf
from spectralclustersupertree_b200 import _app
from spectralclustersupertree_b200.tree import NotCompleted, make_tree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", kat_cases()[:8], ids=lambda c: c["name"])
def test_sc_supertree_app_known_answers(engine, monkeypatch, case):
    """ref: tests/test_app.py:12-27 (scs_test_app)"""
    from spectralclustersupertree_b200 import engine as engine_mod

    monkeypatch.setattr(engine_mod, "_DEFAULT", engine)
    app = _app.get_app("sc_supertree", **case["kwargs"])
    result = app([make_tree(s) for s in case["trees"]]).sorted()
    assert result.same_shape(make_tree(case["expected"]).sorted()), str(result)


def test_load_trees_then_sc_supertree_pipeline(engine, monkeypatch, tmp_path):
    """ref: tests/test_app.py:30-49 (scs_test_pipeline)"""
    from spectralclustersupertree_b200 import engine as engine_mod

    monkeypatch.setattr(engine_mod, "_DEFAULT", engine)
    fixture = json.loads((GOLDEN / "fixture_supertriplets.json").read_text())
    path = tmp_path / "source.tre"
    path.write_text("\n".join(fixture["trees"]) + "\n")
    trees = _app.get_app("load_trees")(str(path))
    assert len(trees) == len(fixture["trees"])
    tree = _app.get_app("sc_supertree", pcg_weighting=fixture["weighting"])(trees)
    assert rf(tree, make_tree(fixture["expected"])) == 0
    with pytest.raises(TypeError, match="Invalid Path Type"):
        _app.get_app("load_trees")(3)


def test_not_completed_trees_are_ignored(engine, monkeypatch):
    from spectralclustersupertree_b200 import engine as engine_mod

    monkeypatch.setattr(engine_mod, "_DEFAULT", engine)
    t1 = "(a,(b,(c,d)))"
    app = _app.get_app("sc_supertree")
    result = app([NotCompleted("ERROR", "test", "upstream"), make_tree(t1)])
    assert result.sorted().same_shape(make_tree(t1).sorted())


def test_outgroup_root():
    """ref: _app.py:64-96"""
    tree = make_tree("((a,b),(c,(d,e)));")
    rooted = _app.get_app("outgroup_root", priority_outgroups=["zz", "d", "a"])(tree)
    assert sorted(rooted.get_tip_names()) == ["a", "b", "c", "d", "e"]
    assert any(child.is_tip() and child.name == "d" for child in rooted)
    with pytest.raises(ValueError, match="does not contain any tip names"):
        _app.get_app("outgroup_root", priority_outgroups=["x", "y"])(tree)
