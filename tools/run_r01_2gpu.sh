#!/bin/bash
# two-GPU validation: the COMM_ALLREDUCE command across two ranks + the data-parallel bench line
timeout 400 python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{"metric"' | tail -1 | tee gpurun_out/r01_bench_dp2.json | cut -c1-600
