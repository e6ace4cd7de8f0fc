#!/bin/bash
for id in 1 2 5 8 10 12 14; do timeout 60 ./build/umma_probe $id 2>&1 | grep -E "FAIL|probe" | cut -c1-200; done
echo "=== timing persistent (8 epilogue warps) ==="
timeout 200 ./build/umma_probe 15 2>&1 | tail -9
timeout 200 ./build/umma_probe 16 2>&1 | tail -13
