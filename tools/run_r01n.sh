#!/bin/bash
for v in 2 1; do
CCV_NNC_SM100_PERSISTENT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PERSISTENT=$v', {k:d[k] for k in ('value','ms_per_step')}, d['per_op']['conv_fwd']['ms'], d['per_op']['conv_bwd']['ms'])"
done
