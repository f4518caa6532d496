"""Build ``libscs_b200.so`` in-tree with nvcc for sm_100a.

    python -m spectralclustersupertree_b200.build [--force]

The library is plain CUDA C++ behind the C ABI of ``include/scs_b200.h``; it has no Python or
torch dependency.  nvcc cross-compiles without a GPU, so this also runs in the CPU-only build
container; the built ``.so`` is git-ignored but travels with the source tree.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libscs_b200.so"

SOURCES = ["context.cu", "pcg.cu", "components.cu", "contract.cu", "spectral.cu", "small.cu", "shard.cu", "driver.cu", "forest.cpp", "newick.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=true",  # contraction is wanted in the matvec; the graph build uses __dmul_rn/__dadd_rn
    "-Xcompiler", "-fPIC,-O2,-Wall,-fopenmp",
    "-shared",
    "-lgomp",
]  # fmt: skip


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).is_file():
        msg = "nvcc not found: the CUDA toolkit is required to build libscs_b200.so"
        raise RuntimeError(msg)
    return nvcc


def needs_build() -> bool:
    if not LIB.is_file():
        return True
    built = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", CSRC / "shard.cuh", CSRC / "forest.hpp", INCLUDE / "scs_b200.h"]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES] + ["-o", str(LIB)]
    proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        msg = f"nvcc failed with exit code {proc.returncode}"
        raise RuntimeError(msg)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
