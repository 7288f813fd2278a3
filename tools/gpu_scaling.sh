#!/bin/bash
# 8-GPU box: weak-scaling bench lines at N = 8, 4, 2, 1 back to back (value leg only), plus the 2-GPU parity tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
nvidia-smi -L > gpurun_out/scaling_box.txt
for N in ${SWEEP:-8 4 2}; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/scaling_bench_$N.log 2>&1
  echo "N=$N t=$(( $(date +%s) - T0 ))"
done
timeout 120 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/scaling_bench_1.log 2>&1
[ -n "$SKIPTEST" ] || timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 150 > gpurun_out/scaling_pytest_multi.log 2>&1; tail -2 gpurun_out/scaling_pytest_multi.log
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    f="gpurun_out/scaling_bench_%d.log"%n
    try:
        l=[x for x in open(f) if x.startswith("{")]
        d=json.loads(l[-1]); r=d["roofline"]
        if n==1: base=d["value"]
        print("N=%d ms/step %.2f value %.3e eff %.3f pred %.2f corr %.2f prep %s"%(n,d["ms_per_step"],d["value"],d["value"]/(n*base) if base else 0,r["predictor"]["ms_per_launch"],r["corrector"]["ms_per_launch"],d["config"].get("prep")))
    except Exception as e:
        print(f,"FAILED",e, open(f).read()[-500:])
PY
echo "t=$(( $(date +%s) - T0 ))"
