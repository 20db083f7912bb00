#!/bin/bash
# Round-end validation on two B200s (gpurun --gpus 2 -- bash scripts/validate_gpu.sh; the r02f run of round 2): whole GPU suite (incl. the 2-GPU bit-identity test), smoke, bench lines of every
# single-GPU configuration + reference arm + N = 2, launch list + --set full capture, compute-sanitizer on 1 and 2 GPUs.
set -u
O=gpurun_out; mkdir -p $O
export WS_SHARDED_LOG_DIR=$PWD/$O
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 | tee $O/r02f_smoke.log
echo "== benches"
timeout 400 python bench.py > $O/r02f_bench_cfg3_n1.json 2> $O/r02f_bench_cfg3_n1.err; tail -c 300 $O/r02f_bench_cfg3_n1.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02f_bench_cfg3_n1_s20.json 2> /dev/null
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02f_bench_ref_n1.json 2> $O/r02f_bench_ref_n1.err
for c in cfg1 cfg2 cfg4; do timeout 300 python bench.py --workload $c --no-cpu-baseline --no-extra > $O/r02f_bench_${c}_n1.json 2> /dev/null; done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02f_bench_cfg3_n2_s20.json 2> /dev/null
python - <<'PY'
import json
for f in ("cfg3_n1", "cfg3_n1_s20", "ref_n1", "cfg1_n1", "cfg2_n1", "cfg4_n1", "cfg3_n2_s20"):
    try:
        d = json.load(open("gpurun_out/r02f_bench_%s.json" % f))
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("ms_per_frame", {}).items() if k not in ("note", "phases_rank0")},
              d["e2e"].get("checksum"), d["e2e"].get("checksum_split_identical"), d["e2e"].get("checksum_matches_n1"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== ncu launch list + full capture of one frame"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/r02f_launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02f_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite|bin_|onesweep|count_kernel|scan_kernel|preprocess" -s 76 -c 19 \
  -o $O/r02f_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --frames-in-flight 1 > $O/r02f_ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== sanitizers (one GPU)"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_kat or golden_frame" > $O/r02f_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/r02f_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "golden_frame or world1" > $O/r02f_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $O/r02f_synccheck.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_frame or sort_kat or edge_cases or sort_matches or deferred or large_rect" > $O/r02f_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02f_memcheck.log
echo "== sanitizers (sharded frame on two GPUs, every rank under the tool)"
for tool in synccheck memcheck; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29872 --no-python \
    compute-sanitizer --tool $tool --error-exitcode 9 python tests/sharded_sync_worker.py > $O/r02f_${tool}_2gpu.log 2>&1; echo "$tool 2-GPU rc=$?" | tee -a $O/r02f_${tool}_2gpu.log
done
for f in racecheck synccheck memcheck synccheck_2gpu memcheck_2gpu; do grep -E "SUMMARY|identical|rc=" $O/r02f_$f.log | tail -4; done
