#!/bin/bash
# first GPU bring-up: each diagnostic case in its own process, bounded by timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/diag.log 2>&1
for c in "one simt halo bf16 1" "one tcgen05 direct bf16 1" "one tcgen05 direct bf16 2" "one tcgen05 halo bf16 1" \
         "one tcgen05 halo bf16 2" "two tcgen05 halo bf16x3 2" "two tcgen05 direct bf16x3 2" "odd tcgen05 halo bf16x3 2 40 48"; do
  echo "=== $c" >> gpurun_out/diag.log
  timeout 180 python tools/diag_tc.py $c >> gpurun_out/diag.log 2>&1
  echo "exit $?" >> gpurun_out/diag.log
done
tail -60 gpurun_out/diag.log
