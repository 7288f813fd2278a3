#!/bin/bash
# quick loop: parity + bench of the product library and of the variants named on the command line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/quick_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_pytest_gpu.log
tail -4 gpurun_out/quick_pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))"
rm -f gpurun_out/quick_bench_*.log
timeout 150 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e > gpurun_out/quick_bench_main.log 2>&1
for v in "$@"; do
  MRG_LIB=$PWD/variants/libmrg_$v.so timeout 150 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e > gpurun_out/quick_bench_v_$v.log 2>&1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/quick_bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-600:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-36s ms/step %.2f pred %.2f corr %.2f clk %s" % (f[11:], d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["clocks"]["sm_mhz"]))
PY
echo "t=$(( $(date +%s) - T0 ))"
