#!/usr/bin/env python
"""Top stall lines (SASS view) of the k-th kernel matching a regex in an .ncu-rep.
usage: ncu_hot.py rep regex [n_lines] [k]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
k = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[k]; e = starts[k + 1] if k + 1 < len(starts) else len(rows)
h = rows[s + 1]
si, ai, ii = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
body = [r for r in rows[s + 2:e] if len(r) > ai]
tot = sum(int(r[ai] or 0) for r in body)
print("kernel:", rows[s][1][:80], "(%d matching launches)" % len(starts), " samples", tot, " warp-instr", sum(int(r[ii] or 0) for r in body))
for j, r in sorted(enumerate(body), key=lambda t: -int(t[1][ai] or 0))[:n]:
    print("%6d %6.2f%%  #%-5d x%-9s %s" % (int(r[ai] or 0), 100.0 * int(r[ai] or 0) / max(tot, 1), j, r[ii], r[si].strip()[:100]))
