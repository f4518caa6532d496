"""TEST INFRASTRUCTURE ONLY -- lets the reference's own ``scs.py`` import without cogent3.

The reference (``/root/reference/src/sc_supertree``) imports ``cogent3`` (ref: scs.py:7-9,
load.py:4), which is neither installed nor installable in this image.  ``install()`` registers
stub modules ``cogent3``, ``cogent3.app.composable`` and ``cogent3.core.tree`` in
``sys.modules`` that expose the handful of names the reference needs, backed by the tree class
in ``spectralclustersupertree_b200.tree`` (the one restatement of cogent3's tree in this repo;
the reference's hot-path functions are duck-typed).

``load_reference()`` then imports the UNMODIFIED reference package from ``SCS_REFERENCE_DIR``
or ``/root/reference`` and returns its ``scs`` module, or None when the reference tree is not
present (it is not on the GPU box; nothing that runs there may need it).  It is used by
``tests/golden/make_golden.py`` to record golden vectors and by the CPU-only tests that
cross-check the oracle against the reference where the reference is mounted.
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path


def install() -> None:
    if "cogent3" in sys.modules and not getattr(sys.modules["cogent3"], "_scs_b200_shim", False):
        return  # a real cogent3 is importable: leave it alone
    from spectralclustersupertree_b200 import tree as _tree

    cogent3 = types.ModuleType("cogent3")
    cogent3._scs_b200_shim = True
    cogent3.PhyloNode = _tree.PhyloNode
    cogent3.make_tree = _tree.make_tree
    cogent3.load_tree = _tree.load_tree
    app = types.ModuleType("cogent3.app")
    composable = types.ModuleType("cogent3.app.composable")
    composable.NotCompleted = _tree.NotCompleted
    core = types.ModuleType("cogent3.core")
    core_tree = types.ModuleType("cogent3.core.tree")
    core_tree.PhyloNode = _tree.PhyloNode
    core_tree.TreeBuilder = _tree.TreeBuilder
    cogent3.app = app
    app.composable = composable
    cogent3.core = core
    core.tree = core_tree
    sys.modules.update(
        {
            "cogent3": cogent3,
            "cogent3.app": app,
            "cogent3.app.composable": composable,
            "cogent3.core": core,
            "cogent3.core.tree": core_tree,
        }
    )


def reference_dir() -> Path | None:
    for cand in (os.environ.get("SCS_REFERENCE_DIR"), "/root/reference"):
        if cand and (Path(cand) / "src" / "sc_supertree" / "scs.py").is_file():
            return Path(cand)
    return None


def load_reference():
    """The reference's ``sc_supertree.scs`` module, imported unmodified, or None if absent."""
    ref = reference_dir()
    if ref is None:
        return None
    install()
    src = str(ref / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    return importlib.import_module("sc_supertree.scs")
