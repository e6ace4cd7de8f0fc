#!/bin/bash
# data-parallel bench on all GPUs of the box (weak scaling, 256 images per GPU, one COMM_ALLREDUCE command per step)
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{"metric"|Error|error' | tail -2 | tee gpurun_out/r01_bench_dp$N.json | cut -c1-700
