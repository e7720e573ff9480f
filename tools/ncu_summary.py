"""Compact per-launch table from an .ncu-rep (full set) or from a gpu__time_duration launch list.

    python tools/ncu_summary.py rep  gpurun_out/x.ncu-rep  > profiles/x.md
    python tools/ncu_summary.py list gpurun_out/launches.csv > profiles/launches.md

`rep`  : one row per profiled launch: duration, DRAM bytes read/written, DRAM %, tensor-pipe %, L2 %, registers.
`list` : per-kernel-name totals and the share of the step each kernel class takes (ncu serialises launches with
         cold caches, so shares are comparable with bench.py's CUDA-event shares, absolutes are not).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("rajni::", "")


def rep(path: str):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    cols = [("time_us", "gpu__time_duration.sum", "time"), ("dram_rd_MB", "dram__bytes_read.sum", None),
            ("dram_wr_MB", "dram__bytes_write.sum", None),
            ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
            ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
            ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
            ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
            ("regs", "launch__registers_per_thread", 1), ("grid", "launch__grid_size", 1)]
    units = rows[1]
    print("| # | kernel | " + " | ".join(c[0] for c in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for n, r in enumerate(rows[2:]):
        out = []
        for label, key, scale in cols:
            if key not in col or r[col[key]] == "":
                out.append("-")
                continue
            v = float(r[col[key]].replace(",", ""))
            if scale == "time":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[col[key]].lower(), 1.0)
                out.append(f"{v:.1f}")
            elif scale is None:                    # bytes with a unit column
                u = units[col[key]].lower()
                v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
                out.append(f"{v:.1f}")
            else:
                v *= scale
                out.append(f"{v:.1f}" if label.endswith("%") or label == "time_us" else f"{v:.0f}")
        print(f"| {n} | {short(r[col['Kernel Name']])} | " + " | ".join(out) + " |")


def launch_list(path: str):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        d = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"].replace(",", "")) * 1e-3
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total:.1f} us serialised (cold-cache) total\n")
    print("| kernel | launches | total_us | share |")
    print("|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us:.1f} | {us / total:.3f} |")


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else launch_list)(sys.argv[2])
