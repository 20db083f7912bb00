#!/bin/bash
# 8-GPU measurement set: cfg3 strong scaling with 2 and 3 frames in flight, and cfg5 (24 M sharded).
for d in 2 3; do
  timeout 240 python bench.py --gpus 8 --steps 144 --frames-in-flight $d > gpurun_out/b8_cfg3_d$d.json 2> gpurun_out/b8_cfg3_d$d.err
done
timeout 300 python bench.py --gpus 8 --steps 72 --workload cfg5 > gpurun_out/b8_cfg5.json 2> gpurun_out/b8_cfg5.err
tail -c 400 gpurun_out/b8_cfg5.err
python - <<'PY'
import json
for f in ("b8_cfg3_d2", "b8_cfg3_d3", "b8_cfg5"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["value"], d["e2e"]["value"], d["config"].get("bands_tile_rows"), d["config"].get("P_sum"))
    except Exception as e:
        print(f, "ERR", e)
PY
