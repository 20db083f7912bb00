#!/bin/bash
# Round-end validation on one B200: GPU test suite, smoke, default bench + the other single-GPU configs,
# ncu launch list of the same command, compute-sanitizer memcheck of one small frame.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/val_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 | tee gpurun_out/val_smoke.log
timeout 300 python bench.py > gpurun_out/val_bench_cfg3.json 2> gpurun_out/val_bench_cfg3.err; tail -c 300 gpurun_out/val_bench_cfg3.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_bench_ref.json 2> gpurun_out/val_bench_ref.err
for c in cfg1 cfg2 cfg4; do
  timeout 300 python bench.py --workload $c --no-cpu-baseline > gpurun_out/val_bench_$c.json 2> gpurun_out/val_bench_$c.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/val_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > gpurun_out/val_ncu.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "golden_frame or sort_kat or edge_cases" > gpurun_out/val_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/val_memcheck.log
tail -5 gpurun_out/val_memcheck.log
python - <<'PY'
import json
for c in ("cfg3", "cfg1", "cfg2", "cfg4", "ref"):
    try:
        d = json.load(open("gpurun_out/val_bench_%s.json" % c))
        print(c, round(d["value"], 2), round(d["e2e"]["value"], 2), d.get("ms_per_frame", {}).get("sort"), d.get("cpu_baseline"))
    except Exception as e:
        print(c, "ERR", e)
PY
