#!/bin/bash
timeout 300 python -m pytest tests/test_parity_sdpa.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
timeout 200 python tools/bench_sdpa.py 2>&1 | tail -2
