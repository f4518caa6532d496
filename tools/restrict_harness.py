"""Time the batched forest restriction on this host (no GPU): python tools/restrict_harness.py [workload] [threads]

Dumps the workload's flat forest to a scratch directory, compiles tools/restrict_harness.cpp against
libscs_b200.so and runs it."""

from __future__ import annotations

import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from spectralclustersupertree_b200 import build  # noqa: E402


def main() -> None:
    workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
    threads = sys.argv[2] if len(sys.argv) > 2 else "8"
    lib = build.build()
    a = bench.make_workload(workload)
    scratch = Path(tempfile.mkdtemp(prefix="scs_restrict_"))
    np.asarray([len(a["node_offsets"]) - 1, len(a["parent"]), len(a["names"])], dtype=np.int64).tofile(scratch / "hdr.bin")
    for name, key, dtype in (("off", "node_offsets", np.int64), ("par", "parent", np.int32), ("len", "length", np.float64),
                             ("sup", "support", np.float64), ("tax", "taxon", np.int32), ("w", "weights", np.float64)):
        a[key].astype(dtype).tofile(scratch / f"{name}.bin")
    source = (ROOT / "tools" / "restrict_harness.cpp").read_text().replace("/tmp/rp/", f"{scratch}/")
    (scratch / "harness.cpp").write_text(source)
    exe = scratch / "harness"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", f"-I{ROOT / 'include'}", f"-I{build.CSRC}",
                    str(scratch / "harness.cpp"), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"], check=True)
    subprocess.run([str(exe), threads], check=True)


if __name__ == "__main__":
    main()
