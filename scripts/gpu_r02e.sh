#!/bin/bash
# round 2, GPU call E (eight B200s): SURVEY 8(e) "G in {1,2,4,8} images bit-identical" at world 4 and 8 (logs kept),
# strong scaling of cfg3 at N = 4, 8 (the driver's --steps 20 and a longer run), cfg5 (24 M over 8 GPUs) with the
# 1-GPU comparison of the checksum view.
set -u
O=gpurun_out; mkdir -p $O
export WS_SHARDED_LOG_DIR=$PWD/$O
nvidia-smi -L | head -8
echo "== bit-identity at world 4 and 8"
timeout 700 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "4 or 8" 2>&1 | tail -8 | tee $O/r02e_pytest_sharded.log
tail -4 $O/sharded_worker_world4.log $O/sharded_worker_world8.log
run() {  # N steps tag [extra args]
  local n=$1 k=$2 tag=$3; shift 3
  timeout ${TMO:-330} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n + k)) bench.py --gpus $n --steps $k --warmup 5 "$@" > $O/r02e_bench_$tag.json 2> $O/r02e_bench_$tag.err
  echo "$tag rc=$?"; tail -c 300 $O/r02e_bench_$tag.err | tr '\n' ' '; echo
}
echo "== cfg3 strong scaling"
run 8 20 cfg3_n8_s20
run 8 108 cfg3_n8
run 4 20 cfg3_n4_s20
run 4 108 cfg3_n4
echo "== cfg5: 24 M Gaussians over 8 GPUs, checksum view verified against one GPU"
TMO=700 run 8 36 cfg5_n8 --workload cfg5
python - <<'PY'
import json
for f in ("cfg3_n8_s20", "cfg3_n8", "cfg3_n4_s20", "cfg3_n4", "cfg5_n8"):
    try:
        d = json.load(open("gpurun_out/r02e_bench_%s.json" % f))
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), d["config"].get("frames_in_flight"), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["ms_per_frame"].items() if k not in ("note", "phases_rank0")},
              d["e2e"].get("checksum"), d["e2e"].get("checksum_n1_same_view"), d["e2e"].get("checksum_matches_n1"))
        print("    phases", {k: round(v, 3) for k, v in d["ms_per_frame"]["phases_rank0"].items()}, "bands", d["config"].get("bands_tile_rows"))
    except Exception as e:
        print(f, "ERR", e)
PY
