#!/bin/bash
# round 2, GPU call L (eight B200s): the driver's scaling commands (--steps 20) on the final build, default frames in flight
set -u
O=gpurun_out; mkdir -p $O
run() {  # N steps tag
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((30300 + $1 * 10)) bench.py --gpus $1 --steps $2 --warmup 5 > $O/r02l_bench_$3.json 2> $O/r02l_bench_$3.err
  echo "$3 rc=$?"
}
run 8 20 cfg3_n8_s20
run 4 20 cfg3_n4_s20
run 2 20 cfg3_n2_s20
python - <<'PY'
import json
for f in ("cfg3_n2_s20", "cfg3_n4_s20", "cfg3_n8_s20"):
    try:
        d = json.load(open("gpurun_out/r02l_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d["config"]["frames_in_flight"], d["e2e"].get("checksum"), d["e2e"].get("checksum_matches_n1"), d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
