#!/bin/bash
# source-level captures of kernels with low tensor-pipe activity inside the step
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'umma_gemm_persistent_kernel<.*0, .*0, 256, 3, 8>' -c 2 -o gpurun_out/r01_ncu_pers_00_256 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_a.log 2>&1
tail -1 gpurun_out/ncu_a.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k umma_wgrad_taps_kernel -c 1 -o gpurun_out/r01_ncu_wgrad_taps64 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
