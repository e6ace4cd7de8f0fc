#!/bin/bash
bash tools/run_probe5.sh 2>&1 | tee gpurun_out/r01_probe5_wgrad_taps.log
python -m pytest tests/test_comm.py tests/test_parity_norm_resample.py tests/test_resnet_parity.py tests/test_parity_contract.py -m gpu -q -s 2>&1 | grep -E "algo|passed|failed|Error|error|node " | cut -c1-600 | head -60
python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01e.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
