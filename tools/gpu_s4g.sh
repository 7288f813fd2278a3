#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 150 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 3 --planes 1 > gpurun_out/s4g_bench_slab3.log 2>&1
grep '^{' gpurun_out/s4g_bench_slab3.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('slab3 ms/step %.2f pred %.2f corr %.2f prep %s'%(d['ms_per_step'], r['predictor']['ms_per_launch'], r['corrector']['ms_per_launch'], d['config'].get('prep')))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4g_launches_slab3.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --slab-of 8 3 --planes 1 > gpurun_out/s4g_ncu.log 2>&1
tail -c 200 gpurun_out/s4g_ncu.log
