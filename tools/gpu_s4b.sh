#!/bin/bash
# session 4, call B: tensor-TMA particle streams (+ ids/keys in the ring), cheap plane tracking, hull-restricted launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/s4b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4b_pytest_gpu.log
tail -5 gpurun_out/s4b_pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 200 python bench.py --steps 8 --warmup 4 --no-cpu > gpurun_out/s4b_bench_main.log 2>&1
timeout 120 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --planes 1 > gpurun_out/s4b_bench_track.log 2>&1
timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 3 --planes 1 > gpurun_out/s4b_bench_slab3_p1.log 2>&1
timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 0 --planes 1 > gpurun_out/s4b_bench_slab0_p1.log 2>&1
timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 3 --planes 0 > gpurun_out/s4b_bench_slab3_p0.log 2>&1
echo "t=$(( $(date +%s) - T0 ))"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/s4b_bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-800:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    e = d.get("e2e") or {}
    print("%-42s ms/step %.2f pred %.2f corr %.2f e2e_ms %s prep %s clk %s" % (f[11:], d["ms_per_step"], r["predictor"]["ms_per_launch"],
          r["corrector"]["ms_per_launch"], e.get("ms_per_step"), d["config"].get("prep"), d["clocks"]["sm_mhz"]))
PY
for k in k_predict_tile k_correct_tile; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 5 -c 1 \
    -o gpurun_out/s4b_prof_$k -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s4b_ncu_$k.log 2>&1
done
echo "t=$(( $(date +%s) - T0 ))"
