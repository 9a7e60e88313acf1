#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --mode augment --steps 50 --warmup 5 > gpurun_out/bench_aug.json 2> gpurun_out/bench_aug.err; echo "bench aug rc=$?"; cut -c1-400 gpurun_out/bench_aug.json; tail -3 gpurun_out/bench_aug.err
timeout 300 python bench.py --mode augment --steps 200 --warmup 5 --size 128 --batch 5 --unroll 4 --no-cpu > gpurun_out/bench_aug128.json 2> gpurun_out/bench_aug128.err; cut -c1-300 gpurun_out/bench_aug128.json; tail -3 gpurun_out/bench_aug128.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_aug.csv python bench.py --mode augment --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list aug rc=$?"
