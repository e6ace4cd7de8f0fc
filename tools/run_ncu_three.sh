#!/bin/bash
# source-level capture of the epilogue-bound expand 1x1 convolution kernel (persistent, BN = 256, 3 stages, 8 epilogue warps)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'umma_gemm_persistent_kernel<\(int\)0, \(int\)0, \(int\)256, \(int\)3, \(int\)8>' -c 1 -o gpurun_out/r01_ncu_pers_00_256 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log | cut -c1-200
ls -la gpurun_out/r01_ncu_pers_00_256.ncu-rep
