#!/bin/bash
# same-box sweep over the preconditioner's polynomial degrees (Chebyshev steps on F, Schur polynomial)
for cfg in "3 2" "4 2" "5 2" "4 3" "5 3"; do
  set -- $cfg
  python bench.py --steps 40 --warmup 6 --cheb $1 --schur-poly $2 --no-secondary --no-parity --no-strong --no-cpu-baseline 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cheb',$1,'poly',$2,'ms/step',round(d['ms_per_step'],4),'its',d['solver']['fgmres_iters_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],4))"
done
