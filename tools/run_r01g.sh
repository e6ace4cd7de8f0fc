#!/bin/bash
# two-stage (atomic-free) BN / colsum reductions, SGD fusion in the bench, first run of the 16-bit flash attention kernel
python -m pytest tests/test_parity_sdpa.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-400
python -m pytest tests/test_parity_feeders.py tests/test_resnet_parity.py tests/test_parity_contract.py -m gpu -q 2>&1 | tail -6 | cut -c1-600
python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01g.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200
timeout 300 python tools/bench_sdpa.py 2>&1 | tail -3
