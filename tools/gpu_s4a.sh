#!/bin/bash
# session 4, call A: parity of the new plumbing, bench (bind + defer + hints), slab-rank emulation of the 8-GPU
# job on one GPU, kernel variants, one ncu --set full capture of each particle kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/s4a_pytest_gpu.log
tail -5 gpurun_out/s4a_pytest_gpu.log
echo "t=$(( $(date +%s) - T0 ))"
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu > gpurun_out/s4a_bench_main.log 2>&1
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --defer 0 > gpurun_out/s4a_bench_sync.log 2>&1
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --planes 1 > gpurun_out/s4a_bench_track.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 3 --planes 1 > gpurun_out/s4a_bench_slab_p1.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 3 --planes 0 > gpurun_out/s4a_bench_slab_p0.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --slab-of 8 0 --planes 1 > gpurun_out/s4a_bench_slab0_p1.log 2>&1
echo "t=$(( $(date +%s) - T0 ))"
for v in ns4 cminb6 cminb4; do
  MRG_LIB=$PWD/variants/libmrg_$v.so timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e > gpurun_out/s4a_bench_v_$v.log 2>&1
done
echo "t=$(( $(date +%s) - T0 ))"
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/s4a_bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-600:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    e = d.get("e2e") or {}
    print("%-42s ms/step %.2f pred %.2f corr %.2f e2e_ms %s h2d %s d2h %s prep %s clk %s" % (f[11:], d["ms_per_step"], r["predictor"]["ms_per_launch"],
          r["corrector"]["ms_per_launch"], e.get("ms_per_step"), e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), d["config"].get("prep"), d["clocks"]["sm_mhz"]))
PY
for k in k_predict_tile k_correct_tile; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 5 -c 1 \
    -o gpurun_out/s4a_prof_$k -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s4a_ncu_$k.log 2>&1
done
echo "t=$(( $(date +%s) - T0 ))"
ls -la gpurun_out | tail -20
