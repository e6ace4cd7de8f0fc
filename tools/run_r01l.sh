#!/bin/bash
timeout 600 python -m pytest tests/test_comm.py tests/test_abi.py -m gpu -q 2>&1 | tail -4 | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','clocks')})"
