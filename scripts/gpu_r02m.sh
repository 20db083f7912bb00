#!/bin/bash
# round 2, GPU call M (one B200): bit-packed saturation map in the far slab's count kernel, 32-bit stage-1 scan
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r02m_pytest_gpu.log
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02m_bench_cfg3.json 2> /dev/null
timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02m_bench_cfg4.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 19 --csv --log-file $O/r02m_launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > /dev/null 2>&1
python - <<'PY'
import json, csv
for f in ("cfg3", "cfg4"):
    d = json.load(open("gpurun_out/r02m_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"}, d["e2e"]["checksum"], d["e2e"]["checksum_split_identical"])
rows = [r for r in csv.reader(open("gpurun_out/r02m_launches_cfg3.csv")) if len(r) > 5 and r[0].isdigit()]
print([(r[4].split("::")[-1].split("(")[0][:18], round(float(r[-1]) / 1000, 1)) for r in rows])
PY
