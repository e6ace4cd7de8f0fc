#!/bin/bash
timeout 300 python -m pytest tests/test_parity_sdpa.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-400
timeout 300 python tools/bench_sdpa.py 2>&1 | tail -3
echo "--- cluster 2"
CCV_NNC_SM100_FMHA_CLUSTER=2 timeout 300 python tools/bench_sdpa.py 2>&1 | tail -3
