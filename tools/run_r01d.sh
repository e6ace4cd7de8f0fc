#!/bin/bash
# validation of the kernel-selection heuristic build: whole-model parity, bench with the per-command table, ncu launch list
python -m pytest tests/test_resnet_parity.py -m gpu -q -x -s 2>&1 | grep -E "algo|passed|failed|Error|error" | cut -c1-900
python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01d.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
