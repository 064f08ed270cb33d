#!/bin/bash
# same-box sweep: size of the projection space (--guess) x raw solutions kept for its rebuild (DNSB_PKEEP)
# usage: tools/gpu_proj_sweep.sh "16 8" "24 8" ...   (extra environment is inherited)
for cfg in "$@"; do
  set -- $cfg
  DNSB_PKEEP=$2 python bench.py --steps 48 --warmup 6 --guess $1 --no-secondary --no-parity --no-strong --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('guess',$1,'keep',$2,'ms/step',round(d['ms_per_step'],4),'its',round(d['solver']['fgmres_iters_per_step'],3),'relres',d['solver']['max_relres'],'e2e_ms',round(d['e2e']['ms_per_step'],4),'e2e_its',round(d['e2e']['fgmres_iters_per_step'],3))"
done
