#!/bin/bash
# parity + bench of several option sets of the product library
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
i=0
while [ $# -gt 0 ]; do
  i=$((i+1))
  timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e $1 > gpurun_out/bench_opt$i.log 2>&1
  echo "opt$i = $1"
  shift
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_opt*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-300:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-40s ms/step %.2f pred %.2f corr %.2f clocks %s" % (f, d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["clocks"]["sm_mhz"]))
PY
