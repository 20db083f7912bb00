#!/bin/bash
# round 2, GPU call L (eight B200s): the driver's scaling commands on the final build (default frames in flight), + cfg5
set -u
O=gpurun_out; mkdir -p $O
run() {  # N steps tag [extra]
  local n=$1 k=$2 tag=$3; shift 3
  timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((30300 + n * 10 + k % 7)) bench.py --gpus $n --steps $k --warmup 5 "$@" > $O/r02l_bench_$tag.json 2> $O/r02l_bench_$tag.err
  echo "$tag rc=$?"
}
run 8 20 cfg3_n8_s20
run 4 20 cfg3_n4_s20
run 2 20 cfg3_n2_s20
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02l_bench_cfg3_n1_s20.json 2> /dev/null; echo "n1 rc=$?"
run 8 108 cfg3_n8
TMO=600 run 8 72 cfg5_n8 --workload cfg5
python - <<'PY'
import json
for f in ("cfg3_n1_s20", "cfg3_n2_s20", "cfg3_n4_s20", "cfg3_n8_s20", "cfg3_n8", "cfg5_n8"):
    try:
        d = json.load(open("gpurun_out/r02l_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d["config"]["frames_in_flight"], d["e2e"].get("checksum"), d["e2e"].get("checksum_matches_n1"))
    except Exception as e:
        print(f, "ERR", e)
PY
