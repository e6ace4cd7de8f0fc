#!/bin/bash
timeout 1200 python -m pytest tests/test_parity_contract.py tests/test_parity_sdpa.py tests/test_resnet_parity.py tests/test_comm.py -m gpu -q 2>&1 | tail -15 | cut -c1-400
timeout 300 python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01i.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
CCV_NNC_SM100_FUSE_CONV_BN=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
timeout 300 python tools/bench_sdpa.py 2>&1 | tail -2
