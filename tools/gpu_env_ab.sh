#!/bin/bash
# same-box A/B of one environment switch of the library: tools/gpu_env_ab.sh NAME v1 v2 ...
name=$1; shift
for v in "$@"; do
  env $name=$v python bench.py --steps 40 --warmup 6 --no-secondary --no-parity --no-strong --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name=$v','ms/step',round(d['ms_per_step'],4),'its',d['solver']['fgmres_iters_per_step'],'relres',d['solver']['max_relres'],'e2e_ms',round(d['e2e']['ms_per_step'],4),'e2e_its',d['e2e']['fgmres_iters_per_step'])"
done
