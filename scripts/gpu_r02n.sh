#!/bin/bash
# round 2, GPU call N (one B200): frames in flight and near-slab share on the final kernels
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra --frames-in-flight 3 > $O/r02n_bench_cfg3_f3.json 2> /dev/null
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02n_bench_cfg3_f1.json 2> /dev/null
WS_SPLIT_NEAR_PCT=20 timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02n_bench_cfg3_near20.json 2> /dev/null
WS_SPLIT_NEAR_PCT=20 timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02n_bench_cfg4_near20.json 2> /dev/null
WS_SPLIT_NEAR_PCT=30 timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02n_bench_cfg3_near30.json 2> /dev/null
python - <<'PY'
import json
for f in ("cfg3_f3", "cfg3_f1", "cfg3_near20", "cfg4_near20", "cfg3_near30"):
    try:
        d = json.load(open("gpurun_out/r02n_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"})
    except Exception as e:
        print(f, "ERR", e)
PY
