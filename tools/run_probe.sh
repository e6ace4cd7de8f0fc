#!/bin/bash
# runs every probe id in its own process (a faulting variant must not poison the rest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for id in "$@"; do
  echo "=== probe $id ==="
  timeout 120 ./build/umma_probe $id 2>&1 | tail -40
  echo "exit=$?"
done
