"""Opcode histogram of every built object (cuobjdump -sass): the evidence that the contractions are tcgen05/TMA code.
usage: python tools/sass_histogram.py > profiles/rN_sass_histogram.md      (after __graft_entry__.build())"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTMACCTL",
         "SYNCS", "LDGSTS", "HMMA", "MUFU.EX2", "FFMA2", "F2FP", "STG.E.ENL2.256", "LDG.E.ENL2.256", "ACQBULK", "SETMAXREG"]
print("# SASS opcode counts per object (`cuobjdump -sass`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync (must be 0).\n")
print("| object | kernels | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for obj in sorted(glob.glob(os.path.join(ROOT, "rajni_vit_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels = len(re.findall(r"^\s*Function :", sass, flags=re.M))
    counts = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                counts[w] += 1
    print(f"| {os.path.basename(obj)} | {kernels} | " + " | ".join(str(counts[w]) for w in WATCH) + " |")
print("\nVariants seen (opcode with modifiers, whole library):\n")
lib = os.path.join(ROOT, "rajni_vit_b200", "csrc", "librajni_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
var = collections.Counter(re.findall(r"\b((?:UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP)[A-Za-z0-9_.]*)", sass))
for k, v in sorted(var.items()):
    print(f"* `{k}` x {v}")
