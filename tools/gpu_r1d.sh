#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_infer.json
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-200 gpurun_out/bench_train.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list train rc=$?"
