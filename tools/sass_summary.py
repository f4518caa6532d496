"""SASS summary of the heavy kernels of libscs_b200.so (no GPU needed): python tools/sass_summary.py > profiles/rNN_sass_summary.txt

For every kernel whose name matches one of the patterns: the code object's architecture, registers / shared memory /
spills (cuobjdump --dump-resource-usage) and counts of the instruction mnemonics that matter for the design claims
(fp64 FMA/ADD/MUL, 128-bit and 64-bit global loads, shared-memory loads/stores/atomics, shuffles, votes, barriers)."""

from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "spectralclustersupertree_b200" / "libscs_b200.so"
PATTERNS = sys.argv[1:] or ["matvec_rows", "pcg_rows_kernel", "pcg_leaf_stairs", "pcg_bucket_sorted", "pcg_mirror_bits",
                            "pcg_degree_rows", "pcg_fetch_transposed", "small_batch_kernel", "med_matvec", "med_tail",
                            "df_tours", "df_mark", "df_write"]
MNEMONICS = ["DFMA", "DADD", "DMUL", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS",
             "ATOMS", "SHFL", "VOTE", "MATCH", "REDUX", "BAR.SYNC", "LDGSTS", "UTMALDG", "HMMA", "DMMA", "UTCHMMA"]


def demangle(names: list[str]) -> dict[str, str]:
    out = subprocess.run(["cu++filt", *names], capture_output=True, text=True, check=False).stdout.splitlines()
    return dict(zip(names, out, strict=False)) if len(out) == len(names) else {n: n for n in names}


def main() -> None:
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", usage)))
    resources = {}
    for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+ STACK:\d+ SHARED:\d+ LOCAL:\d+)", usage):
        resources[m.group(1)] = m.group(2)
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    print(f"{LIB.name}: code objects for {', '.join(arch)} (cuobjdump, CUDA 12.9)")
    wanted = []
    for block in blocks[1:]:
        name = block.split("\n", 1)[0].strip()
        if any(p in name for p in PATTERNS):
            wanted.append((name, block))
    pretty = demangle([n for n, _ in wanted])
    for name, block in wanted:
        ops = Counter()
        total = 0
        for line in block.splitlines():
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            op = m.group(1).replace(".NA", "")  # LDG.E.NA.128 = ld.global.nc.L1::no_allocate.v2.f64
            for key in MNEMONICS:
                if op == key or op.startswith(key + "."):
                    ops[key] += 1
                    break
        short = pretty.get(name, name).replace("scs::<unnamed>::", "").replace("scs::(anonymous namespace)::", "")
        short = re.sub(r">\(.*$", ">", short) if ">(" in short else re.sub(r"\(.*$", "", short)
        counts = ", ".join(f"{k} {v}" for k, v in ops.items() if v)
        print(f"\n{short}\n    {resources.get(name, '?')}; {total} instructions\n    {counts}")


if __name__ == "__main__":
    main()
