#!/bin/bash
timeout 200 ./build/umma_probe 20 2>&1 | grep -E "FAIL|probe" | cut -c1-200
timeout 200 ./build/umma_probe 21 2>&1 | tail -8
timeout 200 ./build/umma_probe 16 2>&1 | tail -13
timeout 900 python -m pytest tests/test_parity_feeders.py tests/test_resnet_parity.py -m gpu -q 2>&1 | tail -8 | cut -c1-400
timeout 300 python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01j.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
