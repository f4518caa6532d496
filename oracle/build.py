"""TEST INFRASTRUCTURE ONLY -- builds oracle/pcg_oracle.c into oracle/_build/libpcg_oracle.so.

The reference is pure Python (no C/C++ sources), so there is nothing to compile into
``oracle/_ref``; the C file here is the oracle's own restatement.  ``-ffp-contract=off`` keeps
the multiply and the add of ``W += value * weight`` separately rounded, as in the reference.
"""

from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def build(force: bool = False) -> Path:
    out_dir = HERE / "_build"
    out_dir.mkdir(exist_ok=True)
    src = HERE / "pcg_oracle.c"
    lib = out_dir / "libpcg_oracle.so"
    if force or not lib.is_file() or lib.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wextra",
               str(src), "-o", str(lib), "-lm"]  # fmt: skip
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force=True))
