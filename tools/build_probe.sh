#!/bin/bash
# builds build/umma_probe (stand-alone kernel probe: correctness vs a double-precision host loop + timing); needs `make -C ccv_b200/csrc` first
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/umma_probe.cu build/obj/sm100_contract.o build/obj/sm100_ffma.o -o build/umma_probe
