"""CPU: the C-ABI library loads and exports every function include/scs_b200.h declares (no compute
calls), and the product package has no route to the oracle."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

from spectralclustersupertree_b200 import _lib, build

ROOT = Path(__file__).resolve().parent.parent


def declared_functions() -> list[str]:
    text = (ROOT / "include" / "scs_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scs_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(str(path))
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in scs_b200.h but not exported"


def test_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_status_strings_and_version():
    lib = _lib.load()
    assert lib.scs_version() >= 100
    assert lib.scs_status_string(0) == b"ok"
    assert b"device" in lib.scs_status_string(_lib.SCS_ERR_NO_DEVICE)
    assert lib.scs_bit_words(33) == 2


def test_product_never_imports_the_oracle():
    for path in (ROOT / "spectralclustersupertree_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    for path in (ROOT / "spectralclustersupertree_b200" / "csrc").iterdir():
        assert "oracle" not in path.read_text().lower(), path
