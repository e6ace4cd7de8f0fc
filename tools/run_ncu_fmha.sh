#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k fmha_fwd_kernel -c 1 -o gpurun_out/r01_ncu_fmha_bf16_config5 python tools/bench_sdpa.py > gpurun_out/ncu_fmha.log 2>&1
tail -2 gpurun_out/ncu_fmha.log | cut -c1-200
ls -la gpurun_out/r01_ncu_fmha_bf16_config5.ncu-rep
