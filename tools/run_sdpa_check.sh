#!/bin/bash
# usage: bash tools/run_sdpa_check.sh <tag>  -- every SDPA parity test, then the configs[4] timing (forward + backward)
T=$1
timeout 600 python -m pytest tests/test_parity_sdpa.py -q -m gpu > gpurun_out/${T}_sdpa.log 2>&1; echo sdpa rc=$?; tail -n 3 gpurun_out/${T}_sdpa.log
timeout 300 python bench.py --workload sdpa_cfg5 > gpurun_out/${T}_sdpa_cfg5.json 2> gpurun_out/${T}_sdpa_cfg5.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/${T}_sdpa_cfg5.json'))['sdpa']
for k in ('full','causal'): print(k, 'fwd ms', d[k]['ms'], 'tflops', d[k]['tflops'], 'bwd ms', d[k]['backward']['ms'], d[k]['backward']['tflops'])
"
