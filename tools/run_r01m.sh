#!/bin/bash
for id in 1 2 5 8 14; do timeout 60 ./build/umma_probe $id 2>&1 | grep -E "FAIL|probe" | cut -c1-200; done
timeout 200 ./build/umma_probe 15 2>&1 | tail -8
timeout 900 python -m pytest tests/test_parity_contract.py tests/test_parity_feeders.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --per-op gpurun_out/per_op_r01m.json --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'])"
CCV_NNC_SM100_TMA_STORE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('TMA_STORE=0', {k:d[k] for k in ('value','ms_per_step')})"
