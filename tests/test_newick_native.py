"""CPU: ``scs_forest_parse_newick`` (Newick text -> flat forest, no node objects) against the Python path
``load_trees`` + ``Forest.from_trees`` on the reference's fixtures, its inline known-answer cases and syntax
corner cases.  Host code only: no GPU is touched."""

from __future__ import annotations

import json

import numpy as np
import pytest

from helpers import GOLDEN, kat_cases
from spectralclustersupertree_b200.engine import Forest
from spectralclustersupertree_b200.load import load_forest, load_trees
from spectralclustersupertree_b200.tree import NewickError, make_tree


def assert_same_forest(native: Forest, lines: list[str]) -> None:
    trees = [make_tree(s.strip()) for s in lines]
    names = sorted({x for t in trees for x in t.get_tip_names()})
    python = Forest.from_trees(trees, [1.0] * len(trees), names)
    try:
        assert list(native.names) == names
        assert native.num_trees == python.num_trees == len(lines)
        for t in range(len(lines)):
            for a, b in zip(native.tree_arrays(t), python.tree_arrays(t), strict=True):
                assert np.array_equal(a, b, equal_nan=True), (t, lines[t][:80])
        assert np.array_equal(native.weights(), python.weights())
    finally:
        python.close()


@pytest.mark.parametrize("name", ["dcm", "dcm_iq", "supertriplets"])
def test_fixture_files_parse_to_the_same_forest(name):
    lines = json.loads((GOLDEN / f"fixture_{name}.json").read_text())["trees"]
    native = Forest.from_newick("\n".join(lines) + "\n")
    try:
        assert_same_forest(native, lines)
    finally:
        native.close()


def test_known_answer_cases_parse_to_the_same_forest():
    for case in kat_cases():
        lines = case["trees"]
        native = Forest.from_newick("\n".join(lines))
        try:
            assert_same_forest(native, lines)
        finally:
            native.close()


def test_syntax_corner_cases():
    lines = [
        "((a:1.5,b:2e-3)90:0.25,(c,'d e':1)inner:3,f)root;",  # support, internal name, quoted label, polytomy
        "  ( (a , b)[a comment [nested]] , ( c , \"it''s\" ) ) ;",  # whitespace, comments, the other quote
        "(a:1,(b:0,c:-0.5)100.0:1E2);",  # zero / negative lengths, float support
        "a;",  # a lone tip
        "((,y),x);",  # an unnamed tip (two of them would be the same taxon twice: rejected)
        "(a,b)",  # no terminating ';'
        "('a''b':+.5,b:5.);",
    ]
    native = Forest.from_newick("\n".join(lines))
    try:
        assert_same_forest(native, lines)
    finally:
        native.close()


@pytest.mark.parametrize(
    "text",
    ["(a,b;", "(a,b));", "(a,b);x", "(a,b)c d;", "(a:x,b);", "(a,'b);", "(a,b)[open;", "", "(a,b);\n\n(c,d);", "a b;"],
)
def test_syntax_errors_are_newick_errors(text):
    lines = text.split("\n") if text else [""]
    with pytest.raises(NewickError):
        for s in lines:
            make_tree(s.strip())
    with pytest.raises(NewickError):
        Forest.from_newick(text if text else "\n")


def test_load_forest_reads_what_load_trees_reads(tmp_path):
    lines = json.loads((GOLDEN / "fixture_supertriplets.json").read_text())["trees"]
    path = tmp_path / "source.tre"
    path.write_text("\n".join(lines) + "\n")
    forest = load_forest(path)
    try:
        assert forest.num_trees == len(load_trees(path)) == len(lines)
        assert_same_forest(forest, lines)
    finally:
        forest.close()


def test_many_lines_use_the_threaded_path():
    rng = np.random.RandomState(3)
    names = [f"t{i}" for i in range(400)]
    lines = []
    for _ in range(600):
        sub = list(rng.choice(names, size=int(rng.randint(2, 60)), replace=False))
        items = [f"{x}:{rng.uniform(0, 2):.5f}" for x in sub]
        while len(items) > 1:
            a, b = items.pop(int(rng.randint(len(items)))), items.pop(int(rng.randint(len(items))))
            items.append(f"({a},{b}){int(rng.randint(50, 101))}:{rng.uniform(0, 2):.5f}")
        lines.append(items[0].rsplit(")", 1)[0] + ");")
    native = Forest.from_newick("\n".join(lines))
    try:
        assert_same_forest(native, lines)
    finally:
        native.close()


def test_random_newick_agrees_with_the_python_parser():
    """Property test: random trees rendered with random spacing, quoting, comments, lengths and supports parse
    to the same forest natively and through ``make_tree``."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    label = st.one_of(
        st.sampled_from(["a", "b", "c", "d", "e", "f", "g", "t12", "x_1", "Homo sapiens", "it's", "7up", "1e3x"]),
        st.text(alphabet="abcXYZ019_-.", min_size=1, max_size=6),
    )
    number = st.one_of(st.integers(0, 1000).map(str), st.floats(0, 50, allow_nan=False).map(lambda v: f"{v:.6g}"),
                       st.sampled_from(["1e-3", "2.5E+1", ".5", "5.", "-0.25"]))  # fmt: skip

    def quote(name: str, how: int) -> str:
        plain = all(ch not in "(),:;[]'\" \t" for ch in name)
        if how == 0 and plain:
            return name
        q = "'" if how != 2 else '"'
        return q + name.replace(q, q + q) + q

    @st.composite
    def newick(draw, depth=0):
        space = draw(st.sampled_from(["", "", " ", "\t", " [c] ", "[x[y]z]"]))
        if depth >= 4 or draw(st.integers(0, 3)) == 0:
            text = quote(draw(label), draw(st.integers(0, 2)))
        else:
            kids = [draw(newick(depth + 1)) for _ in range(draw(st.integers(1, 4)))]
            text = "(" + ",".join(kids) + ")"
            tag = draw(st.integers(0, 3))
            if tag == 1:
                text += draw(number)  # support
            elif tag == 2:
                text += quote("n" + draw(label), draw(st.integers(0, 2)))  # an internal name
        if draw(st.booleans()):
            text += space + ":" + space + draw(number)
        return space + text + space

    @settings(max_examples=150, deadline=None, derandomize=True, database=None)
    @given(st.lists(newick(), min_size=1, max_size=5), st.booleans())
    def check(trees, semicolon):
        lines = [t + (";" if semicolon else "") for t in trees]
        try:
            parsed = [make_tree(s.strip()) for s in lines]
        except NewickError:
            with pytest.raises(NewickError):
                Forest.from_newick("\n".join(lines))
            return
        if any(len(set(t.get_tip_names())) != len(t.get_tip_names()) for t in parsed):
            # the same taxon on two tips of one tree: refused when the forest is made
            with pytest.raises(NewickError, match="more than one tip"):
                Forest.from_newick("\n".join(lines))
            return
        native = Forest.from_newick("\n".join(lines))
        try:
            assert_same_forest(native, lines)
        finally:
            native.close()

    check()


def test_flat_supertree_is_written_as_newick_natively():
    """``scs_flat_tree_newick`` (no node objects) against ``_tree_from_flat(...).get_newick()``, names that need quoting
    included; the text parses back to the same tree."""
    from spectralclustersupertree_b200.scs import _tree_from_flat, flat_newick

    rng = np.random.RandomState(9)
    names = sorted(["a", "b c", "it's", "t(1)", "x:y", "plain_1", "7up", "semi;colon", "comma,name", "z"])
    for _ in range(25):
        count_tips = int(rng.randint(1, len(names) + 1))
        tips = list(rng.permutation(len(names))[:count_tips])
        # random flat tree: parent[i] < i, tips carry taxa
        parent, taxon = [-1], [-1 if count_tips > 1 else int(tips[0])]
        open_internal = [0] if count_tips > 1 else []
        pending = list(tips) if count_tips > 1 else []
        while pending:
            host = int(rng.choice(open_internal))
            if len(pending) > 2 and rng.random_sample() < 0.4:
                parent.append(host)
                taxon.append(-1)
                open_internal.append(len(parent) - 1)
                # an internal node must end up with children: give it two tips right away
                for _k in range(2):
                    parent.append(len(open_internal) and open_internal[-1])
                    taxon.append(int(pending.pop()))
            else:
                parent.append(host)
                taxon.append(int(pending.pop()))
        parent_a, taxon_a = np.array(parent, dtype=np.int32), np.array(taxon, dtype=np.int32)
        # internal nodes without children would read as tips: drop such cases
        has_child = np.zeros(len(parent), dtype=bool)
        has_child[parent_a[1:]] = True
        if ((taxon_a < 0) & ~has_child).any():
            continue
        text = flat_newick(parent_a, taxon_a, names)
        expected = _tree_from_flat(parent_a, taxon_a, names)
        assert text == expected.get_newick()
        assert make_tree(text).clade_sets() == expected.clade_sets()


def test_branch_lengths_parse_like_float():
    """Branch lengths go through ``std::from_chars`` where they are plain decimals and through ``strtod`` otherwise:
    every form must give the double ``make_tree`` (Python's ``float``) gives, bit for bit, or be refused by both."""
    import struct

    rng = np.random.RandomState(3)
    forms = ["0", "-0.0", "1", "007", "1.", ".5", "-.5", "+.5", "+3", "1e3", "1E3", "1e+3", "1e-3", "2.5E+1", "1e308", "1e-308",
             "4.9e-324", "2e-324", "1e400", "-1e400", "1e-400", "0.1", "0.30000000000000004", "123456789012345678901234567890",
             "3.141592653589793238462643383279", "9007199254740993", "1.7976931348623157e308", "inf", "-inf", "Infinity", "nan",
             "1_0", "1_000.5", "1e1_0", "0_1.2_5"]  # fmt: skip
    forms += [repr(float(v)) for v in rng.uniform(0, 1, 200)]
    forms += [repr(float(v)) for v in 10.0 ** rng.uniform(-30, 30, 200)]
    forms += [f"{v:.3f}" for v in rng.uniform(0, 100, 100)]
    lines = [f"(a:{form},b:1);" for form in forms]
    native = Forest.from_newick("\n".join(lines))
    try:
        assert native.num_trees == len(forms)
        for t, form in enumerate(forms):
            got = native.tree_arrays(t)[1][1]
            want = make_tree(lines[t]).children[0].length
            assert struct.pack("<d", got) == struct.pack("<d", want) or (np.isnan(got) and np.isnan(want)), (form, got, want)
    finally:
        native.close()
    for form in ["0x10", "0x1p3", "1e", "e5", "--1", "1..2", "one", "_1", "1_", "1__0", "1_.5", "1._5", "-", "+", "."]:
        line = f"(a:{form},b:1);"
        python_refuses = False
        try:
            make_tree(line)
        except NewickError:
            python_refuses = True
        assert python_refuses, form
        with pytest.raises(NewickError):
            Forest.from_newick(line + "\n")
