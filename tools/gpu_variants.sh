#!/bin/bash
# usage: tools/gpu_variants.sh v1 v2 ... : bench.py step time for each build/variants/lib_<v>.so
cp dolfin_navier_scipy_b200/libdnsb200.so /tmp/lib_keep.so
for v in "$@"; do
  cp build/variants/lib_$v.so dolfin_navier_scipy_b200/libdnsb200.so
  python bench.py --no-cpu-baseline --steps 40 --warmup 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$v', 'ms/step %.4f'%d['ms_per_step'], 'its %.2f'%d['solver']['fgmres_iters_per_step'], r['kernel'], 'us %.1f'%r['mean_us'], 'frac %.3f'%r['frac'])"
done
cp /tmp/lib_keep.so dolfin_navier_scipy_b200/libdnsb200.so
