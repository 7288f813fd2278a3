#!/bin/bash
# BASELINE configs[3]: weak-scaling sweep 256x128x128 per GPU, 100 ppc at N = 1, 2, 4 (N = 8 is in gpu_multi8.sh)
tag=${1:-r02c4}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P=$((29500 + RANDOM % 400))
timeout 300 python bench.py --config 4 --gpus 1 --steps 4 --warmup 2 --no-e2e --no-cpu > gpurun_out/${tag}_n1.log 2>&1
for N in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) bench.py --gpus $N --config 4 --steps 4 --warmup 2 --no-e2e --no-rank-parity --no-reference-partition > gpurun_out/${tag}_n$N.log 2>&1
done
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_n*.log" % sys.argv[1])):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-1200:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-28s N=%d grid %s ms/step %.2f value %.2f G/s pred %.3f corr %.3f parity %s" % (f[11:], d["n_gpus"], d["config"]["grid"], d["ms_per_step"], d["value"]/1e9, r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["parity"].get("ok")))
PY
