#!/bin/bash
# two-GPU run (gpurun --gpus 2): batch-sharded inference, data-parallel training with the NCCL gradient all-reduce,
# sharded post-processing
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 infer rc=$?"; cut -c1-200 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
timeout 900 $TR bench.py --gpus 2 --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "n2 train rc=$?"; cut -c1-200 gpurun_out/bench_train_n2.json; tail -2 gpurun_out/bench_train_n2.err
timeout 600 $TR bench.py --gpus 2 --mode postprocess --steps 400 --warmup 5 > gpurun_out/bench_post_n2.json 2> gpurun_out/bench_post_n2.err; echo "n2 post rc=$?"; cut -c1-200 gpurun_out/bench_post_n2.json; tail -2 gpurun_out/bench_post_n2.err
