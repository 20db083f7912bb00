#!/usr/bin/env python
"""Per-launch table of an `ncu --set full` report (one frame of bench.py): duration, issue-slot utilisation, pipe
utilisation (FMA / ALU / XU = MUFU / LSU / uniform), shared-memory wavefront pipe, DRAM bytes and throughput, occupancy
limiter -- the numbers DESIGN.md and bench.py's `composite_pipes` quote.
usage: ncu_summary.py report.ncu-rep [out_prefix]   -> prints markdown; with out_prefix also writes <out_prefix>.json"""
import csv
import io
import json
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum", "time"),
    ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("warp-instr M", "smsp__inst_executed.sum", 1e-6),
    ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    ("uniform %", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", 1),
    ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("DRAM rd MB", "dram__bytes_read.sum", None),
    ("DRAM wr MB", "dram__bytes_write.sum", None),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("regs", "launch__registers_per_thread", 1),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle", "branch_resolving"]


def to_bytes(val, unit):
    v = float(val)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    table = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name.split("::")[-1].split("(")[0]
        rec = {"kernel": short}
        for label, metric, scale in COLS:
            if metric not in idx or r[idx[metric]] == "":
                rec[label] = None
                continue
            if scale == "time":
                u = units[idx[metric]].lower()
                rec[label] = float(r[idx[metric]]) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(u, 1e-3)
            elif scale is None:
                rec[label] = to_bytes(r[idx[metric]], units[idx[metric]]) / 1e6
            else:
                rec[label] = float(r[idx[metric]]) * scale
        st = {}
        for s in STALLS:
            m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if m in idx and r[idx[m]] != "":
                st[s] = float(r[idx[m]])
        rec["stalls_per_issue"] = {k: round(v, 2) for k, v in sorted(st.items(), key=lambda t: -t[1])[:4]}
        table.append(rec)
    labels = ["kernel"] + [c[0] for c in COLS] + ["top stalls (warps stalled per issue)"]
    print("| # | " + " | ".join(labels) + " |")
    print("|" + "---|" * (len(labels) + 1))
    for i, rec in enumerate(table):
        cells = [rec["kernel"]] + [("%.1f" % rec[c[0]]) if rec[c[0]] is not None else "" for c in COLS]
        cells.append(", ".join("%s %.2f" % kv for kv in rec["stalls_per_issue"].items()))
        print("| %d | " % i + " | ".join(cells) + " |")
    if len(sys.argv) > 2:
        json.dump(table, open(sys.argv[2] + ".json", "w"), indent=1)


if __name__ == "__main__":
    main()
