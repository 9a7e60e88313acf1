#!/bin/bash
# N-GPU evidence on one box (gpurun --gpus N -- 'bash tools/gpu_nN.sh N'): C4 (batch-sharded inference, 4 sequences per
# GPU, no collective), C4 training (gradient all-reduce started per block from inside the backward vs one collective after
# it; with / without synchronised BatchNorm), and C5 (1024x1024, T=16, one sequence per GPU).  Round 1 measured N=2 only.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
run() {  # run <seconds> <tag> <bench args...>
  local secs=$1 tag=$2; shift 2
  timeout -k 10 $secs $TR bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "n$N $tag rc=$?"; cut -c1-220 gpurun_out/bench_n${N}_$tag.json; grep -v "^\*\|NCCL version\|^$" gpurun_out/bench_n${N}_$tag.err | tail -2
}
run 600 infer --steps 10 --warmup 3 --no-cpu
run 900 train_overlapped --mode train --steps 5 --warmup 3 --no-cpu --allreduce overlapped
run 900 train_single --mode train --steps 5 --warmup 3 --no-cpu --allreduce single
run 900 train_syncbn --mode train --steps 5 --warmup 3 --no-cpu --sync-bn
run 900 c5_infer --steps 5 --warmup 3 --no-cpu --size 1024 --unroll 16 --batch 1
