#!/bin/bash
timeout 900 python -m pytest tests/test_parity_feeders.py tests/test_resnet_parity.py -m gpu -q 2>&1 | tail -4 | cut -c1-600
timeout 300 python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01h.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1300
