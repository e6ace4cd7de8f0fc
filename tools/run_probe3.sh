#!/bin/bash
mkdir -p gpurun_out
for id in 1 2 3 4 5 7 8 9 10 11 12 14; do echo "=== probe $id (persistent) ==="; timeout 60 ./build/umma_probe $id 2>&1 | tail -8 | cut -c1-230; done
echo "=== timing persistent ==="
timeout 200 ./build/umma_probe 15 2>&1 | tail -9
timeout 200 ./build/umma_probe 16 2>&1 | tail -13
echo "=== timing one-tile-per-CTA (CCV_NNC_SM100_PERSISTENT=0) ==="
CCV_NNC_SM100_PERSISTENT=0 timeout 200 ./build/umma_probe 15 2>&1 | tail -9
CCV_NNC_SM100_PERSISTENT=0 timeout 200 ./build/umma_probe 16 2>&1 | tail -13
