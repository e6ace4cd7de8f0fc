#!/bin/bash
# usage: bash tools/run_ncu_profiles.sh <tag>      (one GPU; every number printed by a run under ncu is discarded)
# (a) DRAM traffic / duration / tensor-pipe activity of every tensor-core contraction launch of one eager bench step -> <tag>_contraction_traffic.json
# (b) gpu__time_duration launch list of a short default bench run (CUDA graphs) -> <tag>_launches.csv
# (c) ncu --set full of the configs[1] convolution kernels (tf32, bf16), the attention forward and the two attention backward kernels
T=$1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,lts__t_bytes.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'umma_' -c 420 -o /tmp/${T}_contractions python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --no-cuda-graph > gpurun_out/${T}_ncu_a.log 2>&1; echo a rc=$?
python tools/ncu_traffic.py /tmp/${T}_contractions.ncu-rep 0 gpurun_out/${T}_contraction_traffic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/${T}_ncu_b.log 2>&1; echo b rc=$?
python tools/ncu_launch_list.py gpurun_out/${T}_launches.csv 5 > gpurun_out/${T}_launch_list_summary.txt 2>&1; head -20 gpurun_out/${T}_launch_list_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma -c 6 -o gpurun_out/${T}_conv_cfg2 python tools/bench_conv.py tf32 bf16 1 > gpurun_out/${T}_ncu_c1.log 2>&1; echo c1 rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha -c 4 -o gpurun_out/${T}_fmha python bench.py --workload sdpa_cfg5 > gpurun_out/${T}_ncu_c2.log 2>&1; echo c2 rc=$?
