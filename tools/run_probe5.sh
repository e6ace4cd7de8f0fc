#!/bin/bash
# taps-along-M filter gradient: correctness, then timing against the per-tap kernel; bn=64 persistent A/B
for id in 20 12 14; do timeout 120 ./build/umma_probe $id 2>&1 | grep -E "FAIL|probe" | cut -c1-200; done
echo "=== timing taps-along-M wgrad (default) ==="
timeout 200 ./build/umma_probe 21 2>&1 | tail -8
echo "=== timing per-tap wgrad (CCV_NNC_SM100_WGRAD_TAPS=0) ==="
CCV_NNC_SM100_WGRAD_TAPS=0 timeout 200 ./build/umma_probe 21 2>&1 | tail -8
echo "=== forced persistent (bn=64 included) ==="
CCV_NNC_SM100_PERSISTENT=2 timeout 200 ./build/umma_probe 21 2>&1 | tail -8
