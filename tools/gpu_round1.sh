#!/bin/bash
# First GPU pass: parity tests, smoke, pipe-rate probes, bench lines.  Logs -> gpurun_out/
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv
  nproc; free -g | head -2; which gfortran mpif90 mpiexec || echo "no fortran/mpi on the GPU box"
} > gpurun_out/box.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 tools/microbench > gpurun_out/microbench.log 2>&1
timeout 600 python bench.py --grid 64 64 64 --ppc 64 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_small.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_full.log
