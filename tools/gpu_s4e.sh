#!/bin/bash
# variants + launch list of one step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/s4d_bench_*.log
for v in "$@"; do
  MRG_LIB=$PWD/variants/libmrg_$v.so timeout 150 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e > gpurun_out/s4d_bench_v_$v.log 2>&1
done
timeout 150 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --group-min 1 > gpurun_out/s4d_bench_gm1.log 2>&1
timeout 150 python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --group-min 4 > gpurun_out/s4d_bench_gm4.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/s4d_bench_*.log")):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-600:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-36s ms/step %.2f pred %.2f corr %.2f clk %s" % (f[11:], d["ms_per_step"], r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"], d["clocks"]["sm_mhz"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4e_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s4e_ncu_launches.log 2>&1
tail -c 300 gpurun_out/s4e_ncu_launches.log
