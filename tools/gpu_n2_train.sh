#!/bin/bash
# two-GPU training: gradient exchange started per block from inside the backward vs one collective after it
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for m in overlapped single; do
timeout 400 $TR bench.py --gpus 2 --mode train --steps 5 --warmup 3 --allreduce $m > gpurun_out/bench_train_n2_$m.json 2> gpurun_out/bench_train_n2_$m.err; echo "n2 train $m rc=$?"; cut -c1-200 gpurun_out/bench_train_n2_$m.json; grep -v "^\*\|NCCL version" gpurun_out/bench_train_n2_$m.err | tail -3
done
