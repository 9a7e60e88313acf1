#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_r1d.json; tail -5 gpurun_out/bench_r1d.err
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-330 gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
