#!/bin/bash
mkdir -p gpurun_out
for id in 2 3 4 5 8 9 10 11 12 14; do echo "=== probe $id (default MN: lbo4096 sbo512 layout1) ==="; timeout 120 ./build/umma_probe $id 2>&1 | tail -12; done
echo "=== MN variants on probe 2 / 3 ==="
for v in "4096 1024 1" "512 4096 1" "4096 512 2" "1024 512 1" "4096 256 1"; do set -- $v; echo "--- lbo=$1 sbo=$2 layout=$3"; CCV_NNC_SM100_MN_LBO=$1 CCV_NNC_SM100_MN_SBO=$2 CCV_NNC_SM100_MN_LAYOUT=$3 timeout 60 ./build/umma_probe 2 2>&1 | tail -2; CCV_NNC_SM100_MN_LBO=$1 CCV_NNC_SM100_MN_SBO=$2 CCV_NNC_SM100_MN_LAYOUT=$3 timeout 60 ./build/umma_probe 3 2>&1 | tail -2; done
echo "=== timing ==="
timeout 300 ./build/umma_probe 15 2>&1 | tail -12
timeout 300 ./build/umma_probe 16 2>&1 | tail -16
echo "=== timing BN=64 forced / 128 forced ==="
CCV_NNC_SM100_BN=128 timeout 300 ./build/umma_probe 16 2>&1 | tail -16
