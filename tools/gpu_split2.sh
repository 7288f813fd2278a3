#!/bin/bash
# N-GPU check of the split launch / early partial push: multi-rank tests, default bench line, the same without the split
N=${1:-2}; tag=${2:-r02s$N}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > gpurun_out/${tag}_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_multi.log
tail -8 gpurun_out/${tag}_pytest_multi.log
echo "t=$(( $(date +%s) - T0 ))"
P=$((29500 + RANDOM % 400))
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > gpurun_out/${tag}_bench_$name.log 2>&1; P=$((P+1)); }
run default --steps 20 --warmup 5
run nosplit --steps 10 --warmup 4 --split-push 0 --no-e2e --no-rank-parity --no-reference-partition
echo "t=$(( $(date +%s) - T0 ))"
python - "$tag" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.log" % sys.argv[1])):
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO RESULT", open(f).read()[-1500:]); continue
    d = json.loads(l[-1]); r = d["roofline"]
    print("%-32s N=%d ms/step %.2f value %.2f G/s pred %.3f corr %.3f | e2e %s" % (f[11:], d["n_gpus"], d["ms_per_step"], d["value"]/1e9, r["predictor"]["ms_per_launch"], r["corrector"]["ms_per_launch"],
          d["e2e"] and "%.2f ms %.2f G/s shared_ok %s" % (d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"].get("shared_arrays_equal_device_moments"))))
    print("   parity", json.dumps(d.get("parity")))
    print("   multi_rank_parity", json.dumps(d.get("multi_rank_parity")))
    print("   reference_partition", json.dumps(d.get("reference_partition"))[:300])
    print("   phases max", json.dumps(d["phases"]["max_over_ranks"]))
    for q, pr in enumerate(d["phases"].get("per_rank") or []):
        print("   rank %d" % q, " ".join("%s{%s}" % (k, ",".join("%s=%.2f" % (a, b) for a, b in v.items())) for k, v in pr["detail"].items()))
    print("   rank_sum", d["config"].get("rank_sum"))
PY
