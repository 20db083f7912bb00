#!/bin/bash
# round 2, GPU call O (one B200): do fewer resident CTAs of the persistent kernels let two frames overlap better?
set -u
O=gpurun_out; mkdir -p $O
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 108 --warmup 5 --no-cpu-baseline --no-extra > $O/r02o_bench_$tag.json 2> /dev/null; }
b base WS_DUMMY=1
b sort2 WS_SORT_CTAS_PER_SM=2
b pre2 WS_PRE_CTAS_PER_SM=2
b sort2pre2 WS_SORT_CTAS_PER_SM=2 WS_PRE_CTAS_PER_SM=2
b sort1pre2 WS_SORT_CTAS_PER_SM=1 WS_PRE_CTAS_PER_SM=2
b sort2pre2bin2 WS_SORT_CTAS_PER_SM=2 WS_PRE_CTAS_PER_SM=2 WS_BIN_CTAS_PER_SM=2
python - <<'PY'
import json
for f in ("base", "sort2", "pre2", "sort2pre2", "sort1pre2", "sort2pre2bin2"):
    try:
        d = json.load(open("gpurun_out/r02o_bench_%s.json" % f)); print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 4) for k, v in d["ms_per_frame"].items() if k != "note"})
    except Exception as e:
        print(f, "ERR", e)
PY
