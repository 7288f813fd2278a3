#!/bin/bash
# lane-pair kernels: parity (+ memcheck on the small cases), bench, ncu captures.   usage: gpu_round5.sh [tile]
T=${1:-3}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
MRG_TEST_TILE=$T timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_tile$T.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_tile$T.log
tail -5 gpurun_out/pytest_gpu_tile$T.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged" > gpurun_out/memcheck.log 2>&1; tail -5 gpurun_out/memcheck.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --tile $T > gpurun_out/bench_tile$T.log 2>&1; tail -c 1200 gpurun_out/bench_tile$T.log
for k in "k_lane<true>" "k_lane<false>"; do
  n=$(echo $k | tr -d '<>')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lane" -s $([ "$k" = "k_lane<true>" ] && echo 3 || echo 5) -c 1 \
    -o gpurun_out/prof_$n -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --tile $T > gpurun_out/ncu_$n.log 2>&1
done
ls -la gpurun_out | head -40
