#!/bin/bash
timeout 200 ./build/umma_probe 20 2>&1 | grep -E "FAIL|probe" | cut -c1-200
timeout 200 ./build/umma_probe 5 2>&1 | grep -E "FAIL|probe" | cut -c1-200
timeout 900 python -m pytest tests/test_parity_contract.py tests/test_parity_sdpa.py tests/test_resnet_parity.py -m gpu -q 2>&1 | tail -6 | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --per-op gpurun_out/per_op_r01k.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700
timeout 300 python tools/bench_sdpa.py 2>&1 | tail -2
