"""Summaries of ncu CSV logs for profiles/ (run where the CSV is; no GPU needed).

    python tools/ncu_summary.py launches <gpu__time_duration csv>      per-kernel totals of a launch list
    python tools/ncu_summary.py traffic <ncu --page raw csv> <kernel regex> [out.json]
                                                                     DRAM bytes / duration of the matching launches
"""

from __future__ import annotations

import csv
import json
import re
import sys
from collections import defaultdict


def rows_of(path: str):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"scs::\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(path: str) -> None:
    total = defaultdict(float)
    count = defaultdict(int)
    for row in rows_of(path):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        value = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = short(row["Kernel Name"])
        total[name] += value * scale
        count[name] += 1
    whole = sum(total.values()) or 1.0
    print("kernel,launches,total_us,share")
    for name in sorted(total, key=lambda k: -total[k]):
        print(f"{name},{count[name]},{total[name]:.1f},{total[name] / whole:.4f}")
    print(f"ALL,{sum(count.values())},{whole:.1f},1.0")


SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def traffic(path: str, pattern: str, out: str | None) -> None:
    """`ncu -i x.ncu-rep --page raw --csv`: one row per launch, one column per metric, units in the second row."""
    with open(path, newline="") as fh:
        table = [row for row in csv.reader(fh) if row]
    header, units = table[0], table[1]
    col = {name: header.index(name) for name in ("Kernel Name", "Grid Size", "dram__bytes_read.sum",
                                                 "dram__bytes_write.sum", "gpu__time_duration.sum")}

    def value(row, name):
        return float(row[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)

    launches_ = [row for row in table[2:] if re.search(pattern, row[col["Kernel Name"]])]
    if not launches_:
        sys.exit("no matching launch")
    best = max(launches_, key=lambda row: value(row, "dram__bytes_read.sum"))
    result = {
        "kernel": short(best[col["Kernel Name"]]), "grid": best[col["Grid Size"]],
        "dram_bytes_per_launch": value(best, "dram__bytes_read.sum") + value(best, "dram__bytes_write.sum"),
        "dram_read_bytes": value(best, "dram__bytes_read.sum"), "dram_write_bytes": value(best, "dram__bytes_write.sum"),
        "duration_us": value(best, "gpu__time_duration.sum"), "launches_in_capture": len(launches_),
        "all_launches": [{"grid": row[col["Grid Size"]], "duration_us": value(row, "gpu__time_duration.sum"),
                          "dram_bytes": value(row, "dram__bytes_read.sum") + value(row, "dram__bytes_write.sum")}
                         for row in launches_],
        "source": path,
    }
    text = json.dumps(result, indent=1)
    print(text)
    if out:
        with open(out, "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        sys.exit(__doc__)
