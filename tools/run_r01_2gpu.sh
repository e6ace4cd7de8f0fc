#!/bin/bash
# two-GPU validation: the COMM_ALLREDUCE command across two ranks + the data-parallel bench line
python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-1500
