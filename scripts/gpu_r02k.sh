#!/bin/bash
# round 2, GPU call K (one B200): finer sweep of the near-slab share + the split tests at the chosen value
set -u
O=gpurun_out; mkdir -p $O
for pct in 15 20 30; do
  WS_SPLIT_NEAR_PCT=$pct timeout 300 python bench.py --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02i_bench_cfg3_near$pct.json 2> /dev/null
  WS_SPLIT_NEAR_PCT=$pct timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02i_bench_cfg4_near$pct.json 2> /dev/null
done
python - <<'PY'
import json
for c in ("cfg3", "cfg4"):
    for pct in (15, 20, 25, 30, 35, 50, 65):
        try:
            d = json.load(open("gpurun_out/r02i_bench_%s_near%d.json" % (c, pct)))
            print(c, pct, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["config"]["P_mean"] / 1e6, 2), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"}, d["e2e"]["checksum"], d["e2e"]["checksum_split_identical"])
        except Exception as e:
            print(c, pct, "ERR", e)
PY
echo "== split tests at 25 %"
WS_SPLIT_NEAR_PCT=25 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_sharded.py -m gpu -q -k "split or full_size or cfg4 or world1" 2>&1 | tail -5
