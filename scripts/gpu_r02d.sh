#!/bin/bash
# round 2, GPU call D (one B200): occupancy A/B of the sort pass, expand-kernel A/B, full suite, launch lists.
set -u
O=gpurun_out; mkdir -p $O
echo "== sort pass compiled for 4 CTAs/SM (64 registers)"
for v in c4w4 c4w16; do timeout 120 profiles/microbench/sort_vs_cub_$v 20 > $O/r02d_sort_$v.jsonl 2>&1; echo "$v rc=$?"; cut -c1-300 $O/r02d_sort_$v.jsonl; done
timeout 120 profiles/microbench/sort_vs_cub 20 > $O/r02d_sort_default.jsonl 2>&1; cut -c1-300 $O/r02d_sort_default.jsonl
echo "== pytest -m gpu"
WS_TEST_CPP_TOOL=1 timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -25 | tee $O/r02d_pytest_gpu.log
echo "== bench: expand kernel A/B"
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg3_expand2.json 2> /dev/null
WS_BIN_EXPAND=1 timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg3_expand1.json 2> /dev/null
WS_ACTIVE_CULL=0 timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg3_nocull.json 2> /dev/null
WS_ACTIVE_CULL=0 timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg4_nocull.json 2> /dev/null
timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg4_expand2.json 2> /dev/null
WS_BIN_EXPAND=1 timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02d_bench_cfg4_expand1.json 2> /dev/null
python - <<'PY'
import json
for f in ("cfg3_expand2", "cfg3_expand1", "cfg3_nocull", "cfg4_expand2", "cfg4_expand1", "cfg4_nocull"):
    try:
        d = json.load(open("gpurun_out/r02d_bench_%s.json" % f))
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"}, d["e2e"].get("checksum"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== launch lists (both expand kernels)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 19 --csv --log-file $O/r02d_launches_expand2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > /dev/null 2>&1
WS_BIN_EXPAND=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 19 --csv --log-file $O/r02d_launches_expand1.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > /dev/null 2>&1
python - <<'PY'
import csv
for v in ("expand2", "expand1"):
    try:
        rows = [r for r in csv.reader(open("gpurun_out/r02d_launches_%s.csv" % v)) if len(r) > 5 and r[0].isdigit()]
        print(v, [(r[4].split("::")[-1].split("(")[0][:22], r[-1]) for r in rows])
    except Exception as e:
        print(v, "ERR", e)
PY
