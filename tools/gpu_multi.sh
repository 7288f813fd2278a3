#!/bin/bash
# N GPUs: sharded parity + weak-scaling bench lines, deferred vs synchronous moment sums
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
nvidia-smi -L > gpurun_out/multi_box_$N.txt
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 200 > gpurun_out/multi_pytest_multi.log 2>&1; tail -3 gpurun_out/multi_pytest_multi.log
echo "t=$(( $(date +%s) - T0 ))"
for d in 1 0; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$d bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e --defer $d > gpurun_out/multi_bench_${N}_defer$d.log 2>&1
  grep '^{' gpurun_out/multi_bench_${N}_defer$d.log | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('N=%d defer=%s ms/step %.2f value %.3e pred %.2f corr %.2f prep %s'%(d['n_gpus'], d['config']['defer'], d['ms_per_step'], d['value'], r['predictor']['ms_per_launch'], r['corrector']['ms_per_launch'], d['config'].get('prep')))"
  tail -2 gpurun_out/multi_bench_${N}_defer$d.log | cut -c1-300
done
echo "t=$(( $(date +%s) - T0 ))"
