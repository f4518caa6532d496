"""cogent3 app plug-ins: ``load_trees``, ``sc_supertree``, ``outgroup_root`` (ref: src/sc_supertree/_app.py:34-96).

Where cogent3 is installed the three callables are ``cogent3.app.composable.define_app`` apps exactly as in the
reference (same names, signatures and citation), registered through the ``cogent3.app`` entry points in
``pyproject.toml``, so ``get_app("sc_supertree", ...)`` resolves to the GPU engine.  cogent3 is not part of this
image; without it the same three names are plain callables with the same signatures and behaviour (including
``NotCompleted`` passing through ``sc_supertree``'s tree list untouched, ref: scs.py:82-94), which is what the tests
here exercise.
"""

from __future__ import annotations

import os
from collections.abc import Sequence
from typing import Literal

import numpy as np

from .load import load_trees as _load_trees
from .scs import construct_supertree as _construct_supertree

try:  # pragma: no cover - cogent3 is absent from the build image
    import cogent3
    from cogent3.app.composable import define_app

    HAVE_COGENT3 = not getattr(cogent3, "_scs_b200_shim", False)
except ImportError:
    HAVE_COGENT3 = False

CITATION = {
    "key": "10.3389/fmolb.2024.1432495",
    "author": ["McArthur, Robert N.", "Zehmakan, Ahad N.", "Charleston, Michael A.", "Lin, Yu", "Huttley, Gavin"],
    "title": "Spectral cluster supertree: fast and statistically robust merging of rooted phylogenetic trees",
    "year": 2024,
    "journal": "Frontiers in Molecular Biosciences",
    "volume": 11,
    "pages": "1432495",
    "doi": "10.3389/fmolb.2024.1432495",
    "url": "https://www.frontiersin.org/journals/molecular-biosciences/articles/10.3389/fmolb.2024.1432495",
}


def _load_trees_impl(source_tree_file):
    # ref: _app.py:39-42
    if not isinstance(source_tree_file, (str, os.PathLike)):
        msg = f"Invalid Path Type: '{type(source_tree_file)}'."
        raise TypeError(msg)
    return _load_trees(source_tree_file)


def _sc_supertree_impl(
    trees,
    weights: Sequence[float] | None = None,
    pcg_weighting: Literal["one", "branch", "depth", "bootstrap"] = "one",
    *,
    contract_edges: bool = True,
    random_state: np.random.RandomState | None = None,
):
    # ref: _app.py:55-61
    return _construct_supertree(trees, weights, pcg_weighting, contract_edges=contract_edges, random_state=random_state)


def _outgroup_root_impl(tree, *, priority_outgroups: Sequence[str]):
    """Outgroup root a tree at the first name of ``priority_outgroups`` that is one of its tips
    (ref: _app.py:64-96); ``ValueError`` if none is."""
    tip_names = set(tree.get_tip_names())
    for name in priority_outgroups:
        if name in tip_names:
            return tree.rooted(name)
    msg = f"Tree does not contain any tip names in: {priority_outgroups}"
    raise ValueError(msg)


if HAVE_COGENT3:  # pragma: no cover
    try:
        from citeable import Article

        _cite = Article(**CITATION)
        _define = define_app(cite=_cite)
    except Exception:  # noqa: BLE001 - citeable missing or an older define_app without `cite`
        _define = define_app

    @_define
    def load_trees(source_tree_file: str | os.PathLike) -> list:
        return _load_trees_impl(source_tree_file)

    @_define
    def sc_supertree(
        trees: list,
        weights: Sequence[float] | None = None,
        pcg_weighting: Literal["one", "branch", "depth", "bootstrap"] = "one",
        *,
        contract_edges: bool = True,
        random_state: np.random.RandomState | None = None,
    ) -> "cogent3.PhyloNode":
        return _sc_supertree_impl(trees, weights, pcg_weighting, contract_edges=contract_edges, random_state=random_state)

    @_define
    def outgroup_root(tree: "cogent3.PhyloNode", *, priority_outgroups: Sequence[str]) -> "cogent3.PhyloNode":
        return _outgroup_root_impl(tree, priority_outgroups=priority_outgroups)

else:

    class _App:
        """The calling convention of a cogent3 app without cogent3: configure with keyword arguments, then call
        with the data (``get_app("sc_supertree", pcg_weighting="depth")(trees)``)."""

        def __init__(self, func, **settings) -> None:
            self._func = func
            self._settings = settings

        def __call__(self, data):
            return self._func(data, **self._settings)

        main = __call__

    def load_trees(**settings) -> _App:
        return _App(_load_trees_impl, **settings)

    def sc_supertree(**settings) -> _App:
        return _App(_sc_supertree_impl, **settings)

    def outgroup_root(**settings) -> _App:
        return _App(_outgroup_root_impl, **settings)


def get_app(name: str, **settings):
    """``cogent3.get_app`` for the three apps of this package (ref: tests/test_app.py:19-24)."""
    if HAVE_COGENT3:  # pragma: no cover
        return cogent3.get_app(name, **settings)
    apps = {"load_trees": load_trees, "sc_supertree": sc_supertree, "outgroup_root": outgroup_root}
    return apps[name](**settings)
