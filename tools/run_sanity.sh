#!/bin/bash
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_parity_feeders.py -m gpu -q -k "transfer or sgd or relu" 2>&1 | tail -1
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
