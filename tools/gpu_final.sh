#!/bin/bash
# Round artefacts: parity, smoke, full bench line (value + roofline + e2e + cpu_baseline), reference arm,
# ncu launch list of one step, ncu --set full of the two particle kernels.   usage: gpu_final.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv; nproc; } > gpurun_out/${tag}_box.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q --timeout 100 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --steps 10 --warmup 4 > gpurun_out/${tag}_bench.log 2>&1; tail -c 700 gpurun_out/${tag}_bench.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1; tail -c 300 gpurun_out/${tag}_bench_reference.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in k_predict_tile k_correct_tile; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 5 -c 1 \
    -o gpurun_out/${tag}_prof_$k -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_$k.log 2>&1
done
echo "t=$(( $(date +%s) - T0 ))"
ls -la gpurun_out | grep $tag
