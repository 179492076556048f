#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py report.ncu-rep blend_backward [min_pct]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = {}
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        i_inst, i_samp = r.index("Instructions Executed"), r.index("# Samples")
    elif hdr and len(r) > i_inst and r[2] == "-":  # a source line row (SASS rows carry an address)
        try:
            key = (fname, int(r[0]))
        except ValueError:
            continue
        a = agg.setdefault(key, [r[1], 0.0, 0.0])
        a[1] += float(r[i_inst] or 0)
        a[2] += float(r[i_samp] or 0)
ti = sum(a[1] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
for (f, ln), a in sorted(agg.items()):
    if 100 * a[1] / ti >= thr or 100 * a[2] / ts >= thr:
        print(f"{f}:{ln:<4d} {100*a[1]/ti:5.1f}% inst {100*a[2]/ts:5.1f}% samp | {a[0].strip()[:100]}")
