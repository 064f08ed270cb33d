#!/bin/bash
# one-line summary of a bench run: tools/gpu_bench_line.sh <label> [bench args...]
label=$1; shift
timeout 300 python bench.py --no-cpu-baseline --no-profile --no-secondary --no-parity "$@" 2>&1 | tail -1 > gpurun_out/bench_$label.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$label.json"))
    print("== $label", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "its", d["solver"]["fgmres_iters_per_step"], "relres %.2e"%d["solver"]["max_relres"], "launches", d["gpu_launches"])
except Exception as ex:
    print("== $label FAILED", ex); print(open("gpurun_out/bench_$label.json").read()[-800:])
PY
