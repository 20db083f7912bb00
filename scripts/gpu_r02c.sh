#!/bin/bash
# round 2, GPU call C (two B200s): new binning kernel + sharded-frame CUDA graph + occlusion split inside bands:
# parity suite, 2-GPU bit-identity worker, bench at N = 1 and 2.
set -u
O=gpurun_out; mkdir -p $O
export WS_SHARDED_LOG_DIR=$PWD/$O
echo "== pytest (parity + sharded + scale)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_scale.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -30 | tee $O/r02c_pytest_gpu.log
echo "== host submission cost (sharded frame now replays a CUDA graph)"
timeout 200 python scripts/host_cost.py > $O/r02c_host_cost.json 2> $O/r02c_host_cost.err; cat $O/r02c_host_cost.json
echo "== bench N=1"
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02c_bench_cfg3_n1.json 2> $O/r02c_bench_cfg3_n1.err; tail -c 300 $O/r02c_bench_cfg3_n1.err
echo "== bench N=2 (108 steps, then the driver's 20)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 108 --warmup 5 > $O/r02c_bench_cfg3_n2.json 2> $O/r02c_bench_cfg3_n2.err; tail -c 600 $O/r02c_bench_cfg3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02c_bench_cfg3_n2_s20.json 2> $O/r02c_bench_cfg3_n2_s20.err
python - <<'PY'
import json
for f in ("cfg3_n1", "cfg3_n2", "cfg3_n2_s20"):
    try:
        d = json.load(open("gpurun_out/r02c_bench_%s.json" % f))
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["ms_per_frame"].items() if k not in ("note", "phases_rank0")},
              d["e2e"].get("checksum"), d["e2e"].get("checksum_n1_same_view"), d["e2e"].get("checksum_matches_n1"), d["e2e"].get("checksum_split_identical"))
        if "phases_rank0" in d["ms_per_frame"]: print("    phases", {k: round(v, 3) for k, v in d["ms_per_frame"]["phases_rank0"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
