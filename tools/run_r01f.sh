#!/bin/bash
# validation of: SGD multi-tensor fusion, vectorised im2col, pooling index math, bn=64 persistent fprop + ncu of the HBM-bound kernels
python -m pytest tests/test_parity_feeders.py tests/test_resnet_parity.py tests/test_parity_contract.py -m gpu -q 2>&1 | tail -4 | cut -c1-600
python bench.py --steps 5 --warmup 3 --per-op gpurun_out/per_op_r01f.json --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_reduce_kernel|bn_apply_vec_kernel' -c 6 -o gpurun_out/r01_ncu_bn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_bn.log 2>&1
tail -2 gpurun_out/ncu_bn.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'umma_wgrad_taps_kernel|colsum_vec_kernel|im2col_kernel|pool_bwd_kernel' -c 8 -o gpurun_out/r01_ncu_misc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/ncu_misc.log 2>&1
tail -2 gpurun_out/ncu_misc.log | cut -c1-200
