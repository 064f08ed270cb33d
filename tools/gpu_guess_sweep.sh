#!/bin/bash
# same-box sweep over the number of recycled solutions (projection pairs) of the ensemble step
for g in 8 12 16 24; do
  python bench.py --steps 40 --warmup 6 --guess $g --no-secondary --no-parity --no-strong --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('guess',$g,'ms/step',round(d['ms_per_step'],4),'its',d['solver']['fgmres_iters_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],4))"
done
