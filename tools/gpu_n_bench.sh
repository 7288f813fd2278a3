#!/bin/bash
N=${1:-2}; tag=${2:-r02x}; shift 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P=$((29500 + RANDOM % 400))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench_slab.log 2>&1
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % sys.argv[1])):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-1500:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-32s N=%d ms/step %.2f value %.2f G/s pred %.3f corr %.3f | e2e %s" % (f[11:], d["n_gpus"], d["ms_per_step"], d["value"]/1e9, r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"],
          d["e2e"] and "%.2f ms %.2f G/s h2d %.0f MB d2h %.0f MB shared_ok %s" % (d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["h2d_bytes_per_step"]/1e6, d["e2e"]["d2h_bytes_per_step"]/1e6, d["e2e"].get("shared_arrays_equal_device_moments"))))
    print("   parity", json.dumps(d.get("parity")))
    print("   multi_rank_parity", json.dumps(d.get("multi_rank_parity")))
    print("   reference_partition", json.dumps(d.get("reference_partition")))
    print("   phases max", json.dumps(d["phases"]["max_over_ranks"]))
    print("   rank_sum", d["config"].get("rank_sum"))
PY
