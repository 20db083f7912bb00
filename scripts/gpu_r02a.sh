#!/bin/bash
# round 2, GPU call A (one B200): sort V2 vs V1 vs CUB, full GPU test suite, bench A/B, --set full capture of one frame,
# host submission cost, sanitizers.  Everything under `timeout`; outputs -> gpurun_out/r02a_*.
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
echo "== sort microbench"; 
for v in 2 1; do WS_SORT_VARIANT=$v timeout 120 profiles/microbench/sort_vs_cub 20 > $O/r02a_sort_v$v.jsonl 2> $O/r02a_sort_v$v.err; echo "variant $v rc=$?"; cat $O/r02a_sort_v$v.jsonl; done
if ! grep -q '"identical_to_cub": true' $O/r02a_sort_v2.jsonl || grep -q '"identical_to_cub": false' $O/r02a_sort_v2.jsonl; then
  echo "SORT V2 FAILED -> falling back to V1 for the rest of this call"; export WS_SORT_VARIANT=1
fi
echo "== pytest -m gpu"
WS_TEST_CPP_TOOL=1 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $O/r02a_pytest_gpu.log
echo "== bench cfg3 (default variant) + V1"
timeout 300 python bench.py --steps 108 --warmup 5 > $O/r02a_bench_cfg3.json 2> $O/r02a_bench_cfg3.err; tail -c 400 $O/r02a_bench_cfg3.err
WS_SORT_VARIANT=1 timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02a_bench_cfg3_sortv1.json 2> $O/r02a_bench_cfg3_sortv1.err
timeout 300 python bench.py --workload cfg4 --steps 72 --warmup 5 --no-cpu-baseline --no-extra > $O/r02a_bench_cfg4.json 2> $O/r02a_bench_cfg4.err
python - <<'PY'
import json
for f in ("r02a_bench_cfg3", "r02a_bench_cfg3_sortv1", "r02a_bench_cfg4"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"}, d["e2e"].get("checksum"), d["e2e"].get("checksum_split_identical"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== host submission cost"
timeout 200 python scripts/host_cost.py > $O/r02a_host_cost.json 2> $O/r02a_host_cost.err; cat $O/r02a_host_cost.json
echo "== ncu launch list + full capture of one frame"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/r02a_launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02a_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite|bin_|onesweep|count_kernel|scan_kernel|preprocess" -s 76 -c 19 \
  -o $O/r02a_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02a_ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== sanitizers"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_kat or golden_frame" > $O/r02a_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/r02a_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "golden_frame or world1" > $O/r02a_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/r02a_synccheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_frame or sort_kat or edge_cases or sort_matches" > $O/r02a_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02a_memcheck.log
tail -3 $O/r02a_racecheck.log $O/r02a_synccheck.log $O/r02a_memcheck.log
ls -la $O | tail -30
