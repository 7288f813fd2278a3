#!/bin/bash
# quick 1-GPU loop: parity + bench of the product library (+ variants named on the command line)
tag=${TAG:-r02b}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
tail -6 gpurun_out/${tag}_pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 200 python bench.py --steps 10 --warmup 4 --no-cpu > gpurun_out/${tag}_bench_main.log 2>&1
for v in "$@"; do
  MRG_LIB=$PWD/variants/libmrg_$v.so timeout 150 python bench.py --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${tag}_bench_v_$v.log 2>&1
done
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % sys.argv[1])):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-800:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-40s ms/step %.2f pred %.3f corr %.3f e2e %s parity %s clk %s" % (f[11:], d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["e2e"] and round(d["e2e"]["ms_per_step"], 2), d["parity"].get("ok"), d["clocks"]["sm_mhz"]))
    print("   parity", json.dumps(d["parity"]))
PY
echo "t=$(( $(date +%s) - T0 ))"
