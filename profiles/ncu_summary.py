#!/usr/bin/env python
"""Per-launch table of an `ncu --set full` report (one frame of bench.py): duration, issue-slot utilisation, pipe
utilisation (FMA / ALU / XU = MUFU / LSU / uniform), shared-memory wavefront pipe, DRAM bytes and throughput, occupancy
limiter -- the numbers DESIGN.md and bench.py's `composite_pipes` quote.
usage: ncu_summary.py report.ncu-rep [out_prefix] [--bench WORKLOAD "source note"]
  prints markdown; with out_prefix also writes <out_prefix>.json; with --bench (a capture of ONE split frame of that
  workload, 19 launches) also rewrites the two files bench.py reads: profiles/kernel_traffic_<workload>_split.json (DRAM
  bytes per launch and per stage) and profiles/composite_pipes_<workload>.json (measured issue / pipe utilisation)."""
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
    if len(sys.argv) > 2 and not sys.argv[2].startswith("--"):
        json.dump(table, open(sys.argv[2] + ".json", "w"), indent=1)
    if "--bench" in sys.argv:
        import os
        i = sys.argv.index("--bench")
        wl, note = sys.argv[i + 1], sys.argv[i + 2]
        here = os.path.dirname(os.path.abspath(__file__))
        per = [{"index": k, "kernel": t["kernel"], "us": round(t["us"], 1), "dram_read_bytes": t["DRAM rd MB"] * 1e6, "dram_write_bytes": t["DRAM wr MB"] * 1e6} for k, t in enumerate(table)]
        tot = lambda rows: sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in rows)
        pick = lambda name: [r for r in per if r["kernel"].startswith(name)]
        comp = pick("composite_kernel"); sweeps = pick("onesweep")
        depth, tile = sweeps[:4], sweeps[4:]
        stage = {
            "preprocess": {"dram_bytes": tot(pick("count_kernel") + pick("scan_kernel") + pick("preprocess_kernel")), "launches": "count + scan + preprocess"},
            "depth_sort_pass": {"dram_bytes": tot(depth) / max(len(depth), 1), "launches": "average of the 4 depth passes"},
            "binning": {"dram_bytes": tot(pick("bin_")), "launches": "count + scan + expand, both slabs"},
            "tile_sort_pass": {"dram_bytes": tot(tile) / 2.0, "launches": "both slabs, per pass position (2 positions)"},
            "composite": {"dram_bytes": tot(comp), "launches": "near slab (state out) + far slab (state in, pixels out)"},
        }
        json.dump({"source": note, "per_launch": per, "per_stage": stage}, open(os.path.join(here, "kernel_traffic_%s_split.json" % wl), "w"), indent=1)
        comps = [t for t in table if t["kernel"].startswith("composite_kernel")]
        near = max(comps, key=lambda t: t["us"])
        share = 100.0 * near["us"] / max(sum(t["us"] for t in comps), 1e-9)
        pipes = {"kernel": near["kernel"] + " (near slab: %.0f %% of the compositor's time)" % share, "us_under_ncu": round(near["us"], 1),
                 "issue_slots_pct_of_peak": round(near["issue %"], 1), "warp_instructions_M": round(near["warp-instr M"], 1),
                 "pipe_fma_pct": round(near["fma %"], 1), "pipe_alu_pct": round(near["alu %"], 1), "pipe_xu_mufu_pct": round(near["xu %"], 1),
                 "pipe_lsu_pct": round(near["lsu %"], 1), "pipe_uniform_pct": round(near["uniform %"], 1),
                 "shared_memory_wavefronts_pct": round(near["smem wavefronts %"], 1), "dram_pct": round(near["DRAM %"], 1),
                 "stalls_per_issue": near["stalls_per_issue"],
                 "reading": "issue-bound: the schedulers issue on >80 % of the cycles; MUFU (ex2) runs at a fifth of its rate, DRAM at a twentieth"}
        json.dump({"source": note, "metrics": pipes}, open(os.path.join(here, "composite_pipes_%s.json" % wl), "w"), indent=1)


if __name__ == "__main__":
    main()
