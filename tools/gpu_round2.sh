#!/bin/bash
# Second GPU pass: tiled kernels -- parity, bench variants, ncu launch list + full captures.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_tile.log 2>&1; tail -c 1500 gpurun_out/bench_tile.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --tile 0 > gpurun_out/bench_notile.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --fused-keys 0 > gpurun_out/bench_nofuse.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 4 --no-cpu --no-e2e --sort-every 2 > gpurun_out/bench_sort2.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1
# ncu: every launch of one step with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
# ncu: full capture of the two particle kernels (second launch of each = electrons)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_predict_tile|k_correct_tile' -s 2 -c 2 \
  -o gpurun_out/prof_tile python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
