"""Minimal rooted-tree class exposing the cogent3 ``PhyloNode`` surface the path touches.

cogent3 is a dependency of the reference (ref: pyproject.toml:22) but is not installable in
this image, so the host side ships its own node class.  Only the surface the reference
actually uses is provided (SURVEY.md section 8b "Tree object type"):

* hot path (ref: src/sc_supertree/scs.py:570,624-631,560-564): iteration over children,
  ``is_tip()``, ``name``, ``length``, ``support``;
* glue (ref: scs.py:98,186,407,447,450,724,744): ``get_tip_names()``,
  ``get_sub_tree(names, ignore_missing=True, as_rooted=True)``, ``iter_nontips``,
  ``get_newick()``, ``make_tree``, ``TreeBuilder.edge_from_edge`` / ``create_edge``;
* tests (ref: tests/test_spectral_cluster_supertree.py:20-27, tests/helpers.py:10-15):
  ``sorted()``, ``same_shape()``, ``str()``, ``write()``, ``load_tree``.

Every traversal is iterative so caterpillar trees with tens of thousands of levels do not hit
the Python recursion limit.
"""

from __future__ import annotations

import os
from collections.abc import Iterable, Iterator
from pathlib import Path

__all__ = ["NotCompleted", "PhyloNode", "TreeBuilder", "load_tree", "make_tree"]


class NotCompleted:
    """Stand-in for ``cogent3.app.composable.NotCompleted`` (ref: scs.py:8,82-94).

    A failed upstream step arrives in the tree list as one of these and is dropped together
    with its weight.  It is falsy, like the cogent3 original.
    """

    def __init__(self, type_: str = "ERROR", origin: str = "", message: str = "", source=None) -> None:
        self.type = type_
        self.origin = origin
        self.message = message
        self.source = source

    def __bool__(self) -> bool:
        return False

    def __repr__(self) -> str:
        return f"NotCompleted(type={self.type}, origin={self.origin}, message={self.message!r})"


class PhyloNode:
    """A node of a rooted tree; a tree is its root node."""

    __slots__ = ("children", "length", "name", "parent", "support")

    def __init__(
        self,
        name: str | None = None,
        children: Iterable["PhyloNode"] | None = None,
        length: float | None = None,
        support: float | None = None,
    ) -> None:
        self.name = name
        self.length = length
        self.support = support
        self.parent: PhyloNode | None = None
        self.children: list[PhyloNode] = []
        if children is not None:
            for child in children:
                self.append(child)

    # -- structure ---------------------------------------------------------------------------
    def append(self, child: "PhyloNode") -> None:
        child.parent = self
        self.children.append(child)

    def __iter__(self) -> Iterator["PhyloNode"]:
        return iter(self.children)

    def __len__(self) -> int:
        return len(self.children)

    def is_tip(self) -> bool:
        return not self.children

    def is_root(self) -> bool:
        return self.parent is None

    # -- traversals --------------------------------------------------------------------------
    def preorder(self, include_self: bool = True) -> Iterator["PhyloNode"]:
        stack = [self] if include_self else list(reversed(self.children))
        while stack:
            node = stack.pop()
            yield node
            if node.children:
                stack.extend(reversed(node.children))

    def postorder(self, include_self: bool = True) -> Iterator["PhyloNode"]:
        out: list[PhyloNode] = []
        stack = [self]
        while stack:
            node = stack.pop()
            out.append(node)
            stack.extend(node.children)
        out.reverse()
        if not include_self:
            out.pop()
        return iter(out)

    def iter_tips(self) -> Iterator["PhyloNode"]:
        for node in self.preorder():
            if not node.children:
                yield node

    def tips(self) -> list["PhyloNode"]:
        return list(self.iter_tips())

    def iter_nontips(self, include_self: bool = False) -> Iterator["PhyloNode"]:
        for node in self.preorder(include_self=include_self):
            if node.children:
                yield node

    def get_tip_names(self) -> list[str]:
        return [node.name for node in self.iter_tips()]

    # -- copies ------------------------------------------------------------------------------
    def _rebuild(self, keep) -> "PhyloNode | None":
        """Bottom-up copy of the tree keeping only tips for which ``keep(tip)`` holds.

        Unary nodes are merged into their single child and the child's length grows by the
        removed node's length when both are known (otherwise it becomes None).  Returns None if
        nothing is kept.  The returned root may itself be a tip (one tip kept).
        """
        built: dict[int, PhyloNode | None] = {}
        for node in self.postorder():
            if not node.children:
                built[id(node)] = (
                    PhyloNode(node.name, None, node.length, node.support) if keep(node) else None
                )
                continue
            kids = [k for k in (built.pop(id(c)) for c in node.children) if k is not None]
            if not kids:
                built[id(node)] = None
            elif len(kids) == 1:
                only = kids[0]
                if node is not self:
                    if node.length is not None and only.length is not None:
                        only.length = node.length + only.length
                    else:
                        only.length = None
                built[id(node)] = only
            else:
                built[id(node)] = PhyloNode(node.name, kids, node.length, node.support)
        return built[id(self)]

    def copy(self) -> "PhyloNode":
        fresh: dict[int, PhyloNode] = {}
        for node in self.postorder():
            fresh[id(node)] = PhyloNode(
                node.name,
                [fresh.pop(id(c)) for c in node.children],
                node.length,
                node.support,
            )
        return fresh[id(self)]

    deepcopy = copy

    def get_sub_tree(
        self,
        name_list: Iterable[str],
        ignore_missing: bool = False,
        tips_only: bool = False,  # noqa: ARG002  (only tip names are ever matched here)
        as_rooted: bool = False,
    ) -> "PhyloNode":
        """The tree induced on the tips named in ``name_list`` (ref: scs.py:447-452).

        Tips not listed are removed, unary nodes collapse into their child (lengths summed), and
        a unary root is replaced by its first branching descendant.  With ``as_rooted=False`` a
        bifurcating root is additionally collapsed to a trifurcation, as cogent3 does for
        unrooted output; the reference always passes ``as_rooted=True``.
        """
        wanted = set(name_list)
        if not ignore_missing:
            missing = wanted.difference(self.get_tip_names())
            if missing:
                msg = f"edges {sorted(missing)} not found in tree"
                raise ValueError(msg)
        sub = self._rebuild(lambda tip: tip.name in wanted)
        if sub is None:
            msg = "no tips of the tree are in name_list"
            raise ValueError(msg)
        sub.parent = None
        sub.length = None
        if sub.children:
            sub.name = "root"
        if not as_rooted and len(sub.children) == 2:
            for i, child in enumerate(sub.children):
                if child.children:
                    rest = sub.children[1 - i]
                    if rest.length is not None and child.length is not None:
                        rest.length = rest.length + child.length
                    grand = child.children
                    sub.children = []
                    for g in grand:
                        sub.append(g)
                    sub.append(rest)
                    break
        return sub

    def rooted(self, edge_name: str) -> "PhyloNode":
        """A copy re-rooted on the branch above the node called ``edge_name`` (what ``outgroup_root`` needs,
        ref: _app.py:88-92): the new root has two children, that node and the rest of the tree; the branch's length,
        if known, is split evenly between them."""
        copy = self.copy()
        target = next((n for n in copy.preorder() if n.name == edge_name), None)
        if target is None or target.parent is None:
            msg = f"cannot root on '{edge_name}'"
            raise ValueError(msg)
        # walk from the target's parent up to the old root, reversing the parent links
        path = []
        node = target.parent
        while node is not None:
            path.append(node)
            node = node.parent
        half = None if target.length is None else target.length / 2
        below, below_length = target, target.length
        path[0].children.remove(target)
        for k, node in enumerate(path):
            up = path[k + 1] if k + 1 < len(path) else None
            own_length = node.length
            if up is not None:
                up.children.remove(node)
                node.children.append(up)
            node.parent = below if below is not target else None
            node.length = below_length if below is not target else half
            below, below_length = node, own_length
        root = PhyloNode("root")
        target.length = half
        root.append(target)
        root.append(path[0])
        for node in path[1:]:
            node.parent = path[path.index(node) - 1]
        # an old root left with a single child is merged into it
        old_root = path[-1]
        if len(old_root.children) == 1 and old_root.parent is not None:
            only = old_root.children[0]
            if only.length is not None and old_root.length is not None:
                only.length = only.length + old_root.length
            else:
                only.length = None
            holder = old_root.parent
            holder.children[holder.children.index(old_root)] = only
            only.parent = holder
        return root

    def sorted(self, sort_order: Iterable[str] | None = None) -> "PhyloNode":
        """A copy with children ordered by the smallest rank among their tips.

        ``sort_order`` lists tip names in the order wanted; default is alphabetical.
        """
        tip_names = self.get_tip_names()
        order = list(sort_order) if sort_order is not None else []
        known = set(order)
        order.extend(sorted(n for n in tip_names if n not in known))
        rank = {name: i for i, name in enumerate(order)}
        key: dict[int, int] = {}
        fresh: dict[int, PhyloNode] = {}
        for node in self.postorder():
            if not node.children:
                key[id(node)] = rank[node.name]
                fresh[id(node)] = PhyloNode(node.name, None, node.length, node.support)
                continue
            kids = sorted(node.children, key=lambda c: key[id(c)])
            key[id(node)] = key[id(kids[0])]
            fresh[id(node)] = PhyloNode(
                node.name, [fresh.pop(id(c)) for c in kids], node.length, node.support
            )
        return fresh[id(self)]

    def same_shape(self, other: "PhyloNode") -> bool:
        """Topology and tip names equal, child order significant (sort both trees first)."""
        stack = [(self, other)]
        while stack:
            a, b = stack.pop()
            if len(a.children) != len(b.children):
                return False
            if not a.children:
                if a.name != b.name:
                    return False
                continue
            stack.extend(zip(a.children, b.children, strict=True))
        return True

    # -- clades (used by the Robinson-Foulds helper and the tests) -----------------------------
    def clade_sets(self) -> set[frozenset[str]]:
        """The set of tip-name sets below every internal node (rooted clades)."""
        below: dict[int, frozenset[str]] = {}
        clades: set[frozenset[str]] = set()
        for node in self.postorder():
            if not node.children:
                below[id(node)] = frozenset((node.name,))
            else:
                merged = frozenset().union(*(below.pop(id(c)) for c in node.children))
                below[id(node)] = merged
                clades.add(merged)
        return clades

    # -- Newick ------------------------------------------------------------------------------
    def get_newick(
        self,
        with_distances: bool = False,
        semicolon: bool = True,
        escape_name: bool = True,
        with_node_names: bool = False,
    ) -> str:
        pieces: dict[int, str] = {}
        for node in self.postorder():
            if node.children:
                text = "(" + ",".join(pieces.pop(id(c)) for c in node.children) + ")"
                if with_node_names and node.name and node is not self:
                    text += _quote(node.name) if escape_name else node.name
            else:
                name = "" if node.name is None else str(node.name)
                text = _quote(name) if escape_name else name
            if with_distances and node.length is not None and node is not self:
                text += f":{node.length!r}"
            pieces[id(node)] = text
        return pieces[id(self)] + (";" if semicolon else "")

    def __str__(self) -> str:
        return self.get_newick(with_distances=True)

    def __repr__(self) -> str:
        text = self.get_newick(with_distances=False)
        if len(text) > 70:
            text = text[:67] + "..."
        return f'Tree("{text}")'

    def write(self, filename: str | os.PathLike, with_distances: bool = True, format: str | None = None) -> None:  # noqa: A002
        """Write the tree as Newick (the only format the ``scs`` CLI produces, ref: cli.py:39)."""
        if format not in (None, "newick", "nwk", "tre", "tree", "txt"):
            msg = f"unsupported tree format '{format}'"
            raise ValueError(msg)
        Path(filename).write_text(self.get_newick(with_distances=with_distances) + "\n")


def _quote(name: str) -> str:
    if name and not any(ch in name for ch in " ()[]':;,\t\n"):
        return name
    if not name:
        return name
    return "'" + name.replace("'", "''") + "'"


class TreeBuilder:
    """The two ``cogent3.core.tree.TreeBuilder`` calls the reference makes (ref: scs.py:407-408,744-745)."""

    def __init__(self, constructor=PhyloNode) -> None:
        self._cls = constructor or PhyloNode

    def edge_from_edge(self, edge, children, params=None):  # noqa: ARG002
        name = "root" if edge is None else edge.name
        length = None if edge is None else edge.length
        support = None if edge is None else edge.support
        return self._cls(name, list(children), length, support)

    def create_edge(self, children, name, params, name_loaded=True, *_ignored):  # noqa: ARG002
        params = params or {}
        return self._cls(name, list(children or []), params.get("length"), params.get("support"))


# ---------------------------------------------------------------------------------------------
# Newick reading
# ---------------------------------------------------------------------------------------------
class NewickError(ValueError):
    """Raised for malformed Newick text."""


_STRUCT = frozenset("(),:;")


def _tokens(text: str) -> Iterator[tuple[str, str]]:
    """Yield (kind, value): kind is one of the structural characters or 'label'."""
    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch in _STRUCT:
            yield ch, ch
            i += 1
        elif ch.isspace():
            i += 1
        elif ch == "[":
            depth = 1
            i += 1
            while i < n and depth:
                depth += text[i] == "["
                depth -= text[i] == "]"
                i += 1
            if depth:
                msg = "unterminated [comment]"
                raise NewickError(msg)
        elif ch in "'\"":
            buf = []
            i += 1
            while True:
                if i >= n:
                    msg = "unterminated quoted label"
                    raise NewickError(msg)
                if text[i] == ch:
                    if i + 1 < n and text[i + 1] == ch:
                        buf.append(ch)
                        i += 2
                        continue
                    i += 1
                    break
                buf.append(text[i])
                i += 1
            yield "label", "".join(buf)
        else:
            j = i
            while j < n and text[j] not in _STRUCT and text[j] not in "[" and not text[j].isspace():
                j += 1
            yield "label", text[i:j]
            i = j


def _as_number(label: str) -> float | None:
    try:
        return float(label)
    except ValueError:
        return None


def make_tree(treestring: str, name_internal: bool = True) -> PhyloNode:
    """Parse one Newick string into a tree.

    Follows what the reference's tests rely on from ``cogent3.make_tree``:
    ``:x`` is a branch length (ref: tests/test_spectral_cluster_supertree.py:186-187); a numeric
    label on an internal node is its support (ref: same file :217-219); unlabeled internal
    nodes are auto-named ``edge.N`` and the root ``root``.
    """
    open_nodes: list[PhyloNode] = []  # internal nodes whose ')' has not been seen yet
    root: PhyloNode | None = None
    cur: PhyloNode | None = None  # completed node that a following label / length applies to
    expect_length = False
    finished = False

    def new_tip(name: str) -> PhyloNode:
        nonlocal root
        tip = PhyloNode(name)
        if open_nodes:
            open_nodes[-1].append(tip)
        elif root is None:
            root = tip
        else:
            msg = "more than one top-level node"
            raise NewickError(msg)
        return tip

    for kind, value in _tokens(treestring):
        if finished:
            msg = "text after the final ';'"
            raise NewickError(msg)
        if kind == "(":
            if cur is not None:
                msg = "'(' directly after a node"
                raise NewickError(msg)
            inner = PhyloNode()
            if open_nodes:
                open_nodes[-1].append(inner)
            elif root is None:
                root = inner
            else:
                msg = "more than one top-level node"
                raise NewickError(msg)
            open_nodes.append(inner)
        elif kind == ",":
            if not open_nodes:
                msg = "',' outside parentheses"
                raise NewickError(msg)
            if cur is None:
                new_tip("")
            cur = None
            expect_length = False
        elif kind == ")":
            if not open_nodes:
                msg = "unbalanced ')'"
                raise NewickError(msg)
            if cur is None:
                new_tip("")
            cur = open_nodes.pop()
            expect_length = False
        elif kind == ":":
            if cur is None:
                cur = new_tip("")
            expect_length = True
        elif kind == ";":
            finished = True
        elif expect_length:
            number = _as_number(value)
            if number is None:
                msg = f"invalid branch length '{value}'"
                raise NewickError(msg)
            cur.length = number
            expect_length = False
        elif cur is None:
            cur = new_tip(value)
        elif cur.children and cur.name is None and cur.support is None:
            number = _as_number(value)
            if number is None:
                cur.name = value
            else:
                cur.support = number
        else:
            msg = f"unexpected label '{value}'"
            raise NewickError(msg)
    if open_nodes:
        msg = "unbalanced '('"
        raise NewickError(msg)
    if root is None:
        msg = "empty Newick string"
        raise NewickError(msg)
    if name_internal:
        counter = 0
        for node in root.preorder():
            if node.children and not node.name:
                if node is root:
                    node.name = "root"
                else:
                    node.name = f"edge.{counter}"
                    counter += 1
    return root


def load_tree(filename: str | os.PathLike) -> PhyloNode:
    """Read a single Newick tree from a file (ref: tests/helpers.py:10-11)."""
    return make_tree(Path(filename).read_text().strip())
