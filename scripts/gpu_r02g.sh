#!/bin/bash
# round 2, GPU call G (two B200s): the driver's --steps 20 at N = 2 after the sampler-skew fix, and 108 steps
set -u
O=gpurun_out; mkdir -p $O
for k in 20 108; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + k)) bench.py --gpus 2 --steps $k --warmup 5 > $O/r02g_bench_cfg3_n2_s$k.json 2> $O/r02g_bench_cfg3_n2_s$k.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > $O/r02g_bench_ref_n2.json 2> $O/r02g_bench_ref_n2.err
python - <<'PY'
import json
for f in ("cfg3_n2_s20", "cfg3_n2_s108", ):
    d = json.load(open("gpurun_out/r02g_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d["e2e"].get("checksum_matches_n1"), d["clocks"])
d = json.load(open("gpurun_out/r02g_bench_ref_n2.json")); print("ref", d["value"], d["cpu_baseline"]["cores"], d["steps"])
PY
