#!/bin/bash
# lane kernels: parity on both kernel families, bench of the product library and of variants/, ncu of the three lane kernels
T=${1:-3}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
MRG_TEST_TILE=$T timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_tile$T.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_tile$T.log
tail -5 gpurun_out/pytest_gpu_tile$T.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged" > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --tile $T > gpurun_out/bench_main.log 2>&1
for v in variants/libmrg_*.so; do
  [ -f "$v" ] || continue
  n=$(basename $v .so)
  MRG_LIB=$PWD/$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --tile $T > gpurun_out/bench_$n.log 2>&1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-300:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-40s ms/step %.2f pred %.2f corr %.2f clocks %s" % (f, d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["clocks"]["sm_mhz"]))
PY
i=0
for k in a b; do
  i=$((i+1))
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lane|quad" -s $((4+2*i)) -c 1 \
    -o gpurun_out/prof_lane$i -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --tile $T > gpurun_out/ncu_lane$i.log 2>&1
done
