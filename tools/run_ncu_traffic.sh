#!/bin/bash
# ncu --set full over every tensor-core contraction launch of the first bench step (eager, no CUDA graph); the report is reduced
# on the box (it is too large to travel): per-step DRAM bytes, duration, tensor-pipe activity per kernel
timeout 2000 ncu --set full --clock-control none -k regex:'umma_' -c 176 -o /tmp/r01_ncu_contractions python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_contractions.log 2>&1
tail -2 gpurun_out/ncu_contractions.log | cut -c1-200
python tools/ncu_traffic.py /tmp/r01_ncu_contractions.ncu-rep 176 gpurun_out/r01_contraction_traffic.json
ncu -i /tmp/r01_ncu_contractions.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['Kernel Name','launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed']
idx=[h.index(w) for w in want if w in h]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i][:90] for i in idx])
" > gpurun_out/r01_ncu_contractions_summary.csv
wc -l gpurun_out/r01_ncu_contractions_summary.csv
