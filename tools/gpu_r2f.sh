#!/bin/bash
# round-2 artefacts on one GPU: sanity of --config 4/5 on small grids, full bench line, reference arm, ncu launch list + full captures
tag=${1:-r02f}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 200 python bench.py --config 5 --grid 64 32 32 --steps 3 --warmup 2 --no-cpu > gpurun_out/${tag}_cfg5_small.log 2>&1; tail -c 400 gpurun_out/${tag}_cfg5_small.log; echo
timeout 200 python bench.py --config 4 --grid 64 32 32 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_cfg4_small.log 2>&1; tail -c 300 gpurun_out/${tag}_cfg4_small.log; echo
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2>&1; tail -c 600 gpurun_out/${tag}_bench.log; echo
echo "t=$(( $(date +%s) - T0 ))"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1; tail -c 300 gpurun_out/${tag}_bench_reference.log; echo
echo "t=$(( $(date +%s) - T0 ))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-check-direct > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in k_predict_tile k_correct_tile; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 5 -c 1 \
    -o gpurun_out/${tag}_prof_$k -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-check-direct > gpurun_out/${tag}_ncu_$k.log 2>&1
done
echo "t=$(( $(date +%s) - T0 ))"
ls -la gpurun_out | grep $tag
