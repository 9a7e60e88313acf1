#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_halo.json 2> gpurun_out/bench_halo.err; echo "bench halo rc=$?"; cat gpurun_out/bench_halo.json; tail -5 gpurun_out/bench_halo.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --a-mode direct > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench direct rc=$?"; cat gpurun_out/bench_direct.json; tail -5 gpurun_out/bench_direct.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches_r1.csv
