#!/bin/bash
# step time / iteration counts of bench.py for each precision of the dense Schur block
for p in "$@"; do python bench.py --no-cpu-baseline --schur-precision $p 2>/dev/null > gpurun_out/b_$p.json; python -c "
import json,sys
d=json.load(open('gpurun_out/b_$p.json')); r=d['roofline']; ds=r.get('dense_schur') or {}
print('$p', 'ms/step %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'its %.2f'%d['solver']['fgmres_iters_per_step'], 'relres %.1e'%d['solver']['last_relres'], ds.get('kernel'), '%.1f us'%ds.get('mean_us',0))"; done
