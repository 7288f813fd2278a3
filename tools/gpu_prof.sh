#!/bin/bash
# ncu full capture of one launch of each particle kernel (electrons, after warm-up).  usage: gpu_prof.sh <tag> [bench args]
tag=$1; shift
[ -n "$MRG_LIB" ] && export MRG_LIB=$PWD/$MRG_LIB
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for k in k_predict k_correct; do
  MRG_LIB=${MRG_LIB:-} timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 \
    -o gpurun_out/prof_${tag}_$k -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/ncu_${tag}_$k.log 2>&1
done
ls -la gpurun_out
