#!/bin/bash
# Text summary of ncu --set full reports (run where ncu is installed): usage ncu_summary.sh <tag> > profiles/<tag>_ncu_summary.txt
tag=$1
cd "$(dirname "$0")/.."
for k in k_predict_tile k_correct_tile; do
  rep=gpurun_out/${tag}_prof_$k.ncu-rep
  [ -f $rep ] || continue
  echo "=== $k  (ncu --set full --clock-control none, config 2: 128^3 x 64 ppc, one species = 134 217 728 particles per launch)"
  ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; u=r[1]; v=r[-1]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','launch__block_size','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__registers_per_thread','lts__t_sector_hit_rate.pct','sm__cycles_active.avg','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio']
for w in want:
    if w in h: print('  %-90s %-16s %s'%(w,u[h.index(w)],v[h.index(w)]))
"
  ncu -i $rep --page source --csv --print-source sass > /tmp/_src_$k.csv 2>/dev/null
  python tools/ncu_smem.py /tmp/_src_$k.csv
  python tools/ncu_mix.py /tmp/_src_$k.csv 2>/dev/null | head -46
  echo "--- most stalled SASS lines"
  python tools/ncu_hot.py /tmp/_src_$k.csv 14
done
