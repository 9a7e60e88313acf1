#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-330 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
