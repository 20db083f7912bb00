#!/bin/bash
# round 2, GPU call J (eight B200s): more sharded frames in flight
set -u
O=gpurun_out; mkdir -p $O
run() {  # N steps depth tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((30100 + $1 * 10 + $3)) bench.py --gpus $1 --steps $2 --warmup 5 --frames-in-flight $3 > $O/r02j_bench_$4.json 2> $O/r02j_bench_$4.err
  echo "$4 rc=$?"
}
run 8 108 6 cfg3_n8_d6
run 8 108 8 cfg3_n8_d8
run 8 20 6 cfg3_n8_d6_s20
run 4 108 4 cfg3_n4_d4
run 4 108 6 cfg3_n4_d6
run 2 108 3 cfg3_n2_d3
run 2 108 4 cfg3_n2_d4
python - <<'PY'
import json
for f in ("cfg3_n8_d6", "cfg3_n8_d8", "cfg3_n8_d6_s20", "cfg3_n4_d4", "cfg3_n4_d6", "cfg3_n2_d3", "cfg3_n2_d4"):
    try:
        d = json.load(open("gpurun_out/r02j_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d["config"]["frames_in_flight"], d["e2e"].get("checksum_matches_n1"))
    except Exception as e:
        print(f, "ERR", e)
PY
