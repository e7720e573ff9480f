"""profiles/ summary of an `ncu --metrics ... --csv` capture of one forward step: per-launch table + per-class totals.
usage: python tools/step_metrics.py gpurun_out/step_metrics.csv > profiles/xxx.md"""
import collections
import csv
import io
import re
import sys


def short(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("rajni::", "")


GEMM_MODE = {"1": "bias", "2": "bias+GELU", "3": "bias+residual(+row stats)", "0": "generic"}
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u, m = r["Metric Unit"], r["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u.replace("second", "s") if u.endswith("second") else u]
    d[m] = v
# a capture longer than one step: keep the LAST complete step (im2col ... head GEMM)
ids = list(byid)
starts = [i for i, k in enumerate(ids) if "im2col16" in byid[k]["name"]]
if len(starts) > 1:
    span = starts[1] - starts[0]
    full = [s for s in starts if s + span <= len(ids)]
    keep = ids[full[-1]:full[-1] + span]
    byid = collections.OrderedDict((k, byid[k]) for k in keep)
T, RD, WR = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
per = ["| # | kernel | grid | time_us | dram_rd_MB | dram_wr_MB | dram_% | l2_% | tensor_% |", "|---|---|---|---|---|---|---|---|---|"]
tot = collections.OrderedDict()
for i, d in enumerate(byid.values()):
    n = short(d["name"])
    per.append(f"| {i} | {n} | {d['grid']} | {d[T]:.1f} | {d[RD]:.1f} | {d[WR]:.1f} | "
               f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', d.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', float('nan'))):.1f} | "
               f"{d['lts__throughput.avg.pct_of_peak_sustained_elapsed']:.1f} | "
               f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', d.get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', float('nan'))):.1f} |")
    cls = n.split("<")[0]
    m = re.match(r"gemm_bf16_kernel<(\d+), (\d), (\d), (\d)>", n)
    if m:
        cls = f"gemm {GEMM_MODE[m.group(3)]}{' +LN fold' if m.group(4) == '1' else ''}"
    t = tot.setdefault(cls, [0, 0.0, 0.0, 0.0])
    t[0] += 1
    t[1] += d[T]
    t[2] += d[RD]
    t[3] += d[WR]
all_t = sum(t[1] for t in tot.values())
print(f"# ncu per-launch metrics of ONE forward step ({len(byid)} launches)\n")
print("Serialised, cold-cache replays under ncu: compare SHARES with bench.py's CUDA-event shares, not absolutes.\n")
print("## per kernel class\n")
print("| class | launches | time_us | share | dram_rd_MB | dram_wr_MB | dram_total_MB |\n|---|---|---|---|---|---|---|")
for k, t in tot.items():
    print(f"| {k} | {t[0]} | {t[1]:.1f} | {t[1] / all_t:.3f} | {t[2]:.1f} | {t[3]:.1f} | {t[2] + t[3]:.1f} |")
g = [t for k, t in tot.items() if k.startswith("gemm")]
print(f"\nAll GEMM launches of the step: {sum(t[0] for t in g)} launches, {sum(t[1] for t in g):.1f} us, "
      f"DRAM traffic {sum(t[2] + t[3] for t in g):.1f} MB ({sum(t[2] + t[3] for t in g) / sum(t[0] for t in g):.1f} MB per launch)\n")
print("## per launch\n")
print("\n".join(per))
