#!/bin/bash
# State-of-the-world pass: parity, bench per kernel family, sort cadence, ncu launch list + full captures.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for t in 1 2; do
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --tile $t > gpurun_out/bench_tile$t.log 2>&1
done
for se in 2 4; do
  timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --tile 1 --sort-every $se > gpurun_out/bench_tile1_sort$se.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tile1.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --tile 1 > gpurun_out/ncu_launches.log 2>&1
for k in k_predict_tile k_correct_tile; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 \
    -o gpurun_out/prof_$k -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --tile 1 > gpurun_out/ncu_$k.log 2>&1
done
for k in k_predict_pair k_correct_pair; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 \
    -o gpurun_out/prof_$k -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --tile 2 > gpurun_out/ncu_$k.log 2>&1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT"); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-40s ms/step %.2f pred %.2f corr %.2f clocks %s" % (f, d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["clocks"]))
PY
ls -la gpurun_out
