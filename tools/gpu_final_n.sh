#!/bin/bash
# final N-GPU lines of the round: the default bench line (e2e + all-rank parity) and the same step without the split launch
N=${1:-8}; tag=${2:-r02z$N}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/${tag}_box.txt; nproc >> gpurun_out/${tag}_box.txt; free -g >> gpurun_out/${tag}_box.txt
P=$((29500 + RANDOM % 400))
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > gpurun_out/${tag}_bench_$name.log 2>&1; P=$((P+1)); }
run default --steps 20 --warmup 5 --no-reference-partition
[ -z "$NOSPLIT_OFF" ] && run nosplit --steps 20 --warmup 5 --split-push 0 --no-e2e --no-rank-parity --no-reference-partition
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % sys.argv[1])):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-1500:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-32s N=%d ms/step %.2f value %.2f G/s pred %.3f corr %.3f | e2e %s" % (f[11:], d["n_gpus"], d["ms_per_step"], d["value"]/1e9, r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"],
          d["e2e"] and "%.2f ms %.2f G/s h2d %.0f MB d2h %.0f MB shared_ok %s" % (d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["h2d_bytes_per_step"]/1e6, d["e2e"]["d2h_bytes_per_step"]/1e6, d["e2e"].get("shared_arrays_equal_device_moments"))))
    print("   parity", (d.get("parity") or {}).get("ok"), "mrp", json.dumps(d.get("multi_rank_parity")))
    print("   phases max", json.dumps(d["phases"]["max_over_ranks"]))
    for q, pr in enumerate(d["phases"].get("per_rank") or []):
        print("   rank %d" % q, " ".join("%s{%s}" % (k, ",".join("%s=%.2f" % (a, b) for a, b in v.items())) for k, v in pr["detail"].items()))
    print("   rank_sum", d["config"].get("rank_sum"))
PY
