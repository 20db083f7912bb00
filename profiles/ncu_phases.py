#!/usr/bin/env python
"""Program-order stall profile of a kernel: cumulative samples between BAR.SYNC instructions.
usage: ncu_phases.py rep regex [k]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[k]; e = starts[k + 1] if k + 1 < len(starts) else len(rows)
h = rows[s + 1]
si, ai, ii = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
body = [r for r in rows[s + 2:e] if len(r) > ai]
tot = sum(int(r[ai] or 0) for r in body)
print("kernel:", rows[s][1][:70], "samples", tot)
acc = 0; acci = 0; first = 0; marks = ("BAR.SYNC", "MATCH", "WARPSYNC", "SYNCS", "UBLKCP", "ATOM", "RED.", "MEMBAR")
for j, r in enumerate(body):
    acc += int(r[ai] or 0); acci += int(r[ii] or 0)
    src = r[si].strip()
    if "BAR.SYNC" in src or "BAR.RED" in src or j == len(body) - 1:
        print("  lines %4d-%4d  samples %6d (%5.1f%%)  warp-instr %10d   ends: %s" % (first, j, acc, 100.0 * acc / max(tot, 1), acci, src[:60]))
        acc = 0; acci = 0; first = j + 1
