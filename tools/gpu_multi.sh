#!/bin/bash
# multi-GPU: parity of the sharded path + weak-scaling bench lines.   usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/multi_box.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/multi_pytest.log 2>&1; tail -3 gpurun_out/multi_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu > gpurun_out/multi_bench_$N.log 2>&1
grep '^{' gpurun_out/multi_bench_$N.log | tail -1 | cut -c1-400; tail -3 gpurun_out/multi_bench_$N.log | cut -c1-300
