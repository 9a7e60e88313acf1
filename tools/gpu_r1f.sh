#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_augment.py -m gpu -x -q 2>&1 | tail -5
for c in 2 1; do LU_WGRAD_CLUSTER=$c timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train_c$c.json 2> gpurun_out/bench_train_c$c.err; echo "bench train cluster=$c rc=$?"; cut -c1-200 gpurun_out/bench_train_c$c.json; tail -2 gpurun_out/bench_train_c$c.err; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list train rc=$?"
timeout 300 python bench.py --mode augment --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_aug.json 2> gpurun_out/bench_aug.err; cut -c1-200 gpurun_out/bench_aug.json
