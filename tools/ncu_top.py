"""Summarise an .ncu-rep: per kernel key metrics + top stalled SASS lines.  usage: ncu_top.py rep [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"  {h:80s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
for blk in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
    name = lines[0][1][:90]
    h = lines[1]
    data = [r for r in lines[2:] if len(r) == len(h)]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp] or 0) for r in data) or 1
    print("=" * 100)
    print(name, "samples", tot)
    agg = {}
    for r in data:
        for i in stalls:
            agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i] or 0)
    print("  stall mix:", ", ".join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn]:
        st = sorted(((h[i][6:], int(r[i] or 0)) for i in stalls if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:2]
        print(f"  {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>9s} {r[isrc].strip()[:80]:80s} {st}")
