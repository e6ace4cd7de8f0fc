#!/bin/bash
# full GPU suite + smoke + bench (default flags, with the CPU baseline) + reference arm + ncu launch list + fmha bench
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-500
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --per-op gpurun_out/r01_per_op_final.json 2>&1 | tail -1 > gpurun_out/r01_bench_line.json; cut -c1-600 gpurun_out/r01_bench_line.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r01_bench_reference_line.json; cut -c1-400 gpurun_out/r01_bench_reference_line.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/bench_sdpa.py 2>&1 | tail -2 | tee gpurun_out/r01_sdpa_config5.jsonl
