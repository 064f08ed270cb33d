#!/bin/bash
# usage: tools/gpu_env_ab.sh VAR v1 v2 ... : bench.py step time with VAR set to each value
var=$1; shift
for v in "$@"; do env $var=$v python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/b_ab.json; python -c "
import json
d=json.load(open('gpurun_out/b_ab.json')); r=d['roofline']
print('$var=$v', 'ms/step %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'its %.2f'%d['solver']['fgmres_iters_per_step'], 'relres %.2e'%d['solver']['last_relres'])
for k in r['kernel_shares']: print('   %-42s %4d %7.1f us %.3f'%(k['kernel'][:42],k['launches'],k['mean_us'],k['share']))"; done
