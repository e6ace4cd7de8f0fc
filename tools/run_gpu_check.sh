#!/bin/bash
# usage: bash tools/run_gpu_check.sh <tag>   -- contraction parity tests, configs[1] per-mode timing, configs[4] timing, headline bench
T=$1
timeout 900 python -m pytest tests/test_parity_contract.py tests/test_parity_16bit.py tests/test_parity_sdpa.py tests/test_resnet_parity.py -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo pytest rc=$?; tail -n 4 gpurun_out/${T}_pytest.log
timeout 300 python tools/bench_conv.py tf32 3xtf32 bf16 > gpurun_out/${T}_conv_cfg2.jsonl 2>&1; cat gpurun_out/${T}_conv_cfg2.jsonl | cut -c1-300
timeout 300 python bench.py --workload sdpa_cfg5 > gpurun_out/${T}_sdpa_cfg5.json 2> gpurun_out/${T}_sdpa_cfg5.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_sdpa_cfg5.json'))['sdpa']
for k in ('full','causal'): print(k, 'fwd ms', d[k]['ms'], 'tflops', d[k]['tflops'], 'bwd ms', d[k]['backward']['ms'], d[k]['backward']['tflops'])
"
timeout 600 python bench.py --no-variants --no-cpu-baseline --per-op gpurun_out/${T}_f32_per_op.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc=$?; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('f32', d['ms_per_step'], d['value'], d['roofline']['frac'])"
timeout 600 python bench.py --dtype bf16 --no-variants --no-cpu-baseline --per-op gpurun_out/${T}_bf16_per_op.json > gpurun_out/${T}_bench_bf16.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_bf16.json')); print('bf16', d['ms_per_step'], d['value'], d['roofline']['frac'])"
