#!/bin/bash
# round 2, GPU call B (one B200): look-back width A/B, full GPU suite on the fmaf/reciprocal stage 1 + branch-free
# compositor + wide look-back, benches of all single-GPU configs, launch list + --set full capture, sanitizers.
set -u
O=gpurun_out; mkdir -p $O
echo "== look-back width A/B (V1 pass)"
for w in 4 8; do timeout 120 profiles/microbench/sort_vs_cub_w$w 20 > $O/r02b_sort_lb$w.jsonl 2>&1; echo "LB width $w rc=$?"; cut -c1-330 $O/r02b_sort_lb$w.jsonl; done
timeout 120 profiles/microbench/sort_vs_cub 20 > $O/r02b_sort_lb16.jsonl 2>&1; echo "LB width 16 (library) rc=$?"; cut -c1-330 $O/r02b_sort_lb16.jsonl
echo "== pytest -m gpu"
WS_TEST_CPP_TOOL=1 timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $O/r02b_pytest_gpu.log
echo "== benches"
timeout 300 python bench.py --steps 108 --warmup 5 > $O/r02b_bench_cfg3.json 2> $O/r02b_bench_cfg3.err; tail -c 400 $O/r02b_bench_cfg3.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/r02b_bench_cfg3_s20.json 2> /dev/null
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02b_bench_cfg3_f1.json 2> /dev/null
timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra --frames-in-flight 3 > $O/r02b_bench_cfg3_f3.json 2> /dev/null
for c in cfg1 cfg2 cfg4; do timeout 300 python bench.py --workload $c --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02b_bench_$c.json 2> $O/r02b_bench_$c.err; done
python - <<'PY'
import json
for f in ("cfg3", "cfg3_s20", "cfg3_f1", "cfg3_f3", "cfg1", "cfg2", "cfg4"):
    try:
        d = json.load(open("gpurun_out/r02b_bench_%s.json" % f))
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"}, d["e2e"].get("checksum"), d["e2e"].get("checksum_split_identical"),
              "sb_frac", round(d["roofline"]["sort_plus_blend"]["frac"], 3), round(d["roofline"]["sort_plus_blend"]["frac_full_pairs"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== ncu launch list + full capture of one frame"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/r02b_launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02b_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite|bin_|onesweep|count_kernel|scan_kernel|preprocess" -s 76 -c 19 \
  -o $O/r02b_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02b_ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== sanitizers"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_kat or golden_frame" > $O/r02b_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/r02b_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "golden_frame or world1" > $O/r02b_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/r02b_synccheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_frame or sort_kat or edge_cases or sort_matches or deferred" > $O/r02b_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02b_memcheck.log
for f in racecheck synccheck memcheck; do tail -n 4 $O/r02b_$f.log; done
