#!/bin/bash
# timing experiments on k_cheb_step_tile (results are WRONG for dbg != 0): per-kernel event times
for dbg in 0 1 2 3 4 7; do
  DNSB_TILE_DBG=$dbg timeout 200 python bench.py --steps 10 --warmup 3 --spinup 0 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 > gpurun_out/bench_dbg$dbg.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_dbg$dbg.json"))
    print("dbg $dbg", {k:v for k,v in d["kernels"].items() if "tile" in k})
except Exception as ex:
    print("dbg $dbg FAILED", ex)
PY
done
