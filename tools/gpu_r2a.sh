#!/bin/bash
# round-2 first pass: parity (incl. reference-output goldens), smoke, bench line with parity/fp64 objects, reference arm
# (the reference's own code), full-size verify at BASELINE configs[1].   usage: gpu_r2a.sh <tag>
tag=${1:-r02a}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv; nproc; free -g; } > gpurun_out/${tag}_box.txt 2>&1
timeout 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --steps 10 --warmup 4 > gpurun_out/${tag}_bench.log 2>&1; tail -c 1500 gpurun_out/${tag}_bench.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1; tail -c 600 gpurun_out/${tag}_bench_reference.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 600 python bench.py --verify > gpurun_out/${tag}_verify.log 2>&1; tail -c 1200 gpurun_out/${tag}_verify.log
echo "t=$(( $(date +%s) - T0 ))"
