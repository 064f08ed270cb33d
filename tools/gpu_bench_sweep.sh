#!/bin/bash
# usage: tools/gpu_bench_sweep.sh "<label>" "<env assignments>" [bench args...]
label=$1; shift
envs=$1; shift
env $envs timeout 300 python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/bench_$label.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$label.json"))
print("== $label", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "its", d["solver"]["fgmres_iters_per_step"], "relres %.2e"%d["solver"]["last_relres"], "launches", d["gpu_launches"])
for k,v in d.get("kernels",{}).items(): print("   %-44s x%-6s %8.2f us"%(k, v[0], v[1]))
PY
