#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; cat gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 110 -c 1 -o gpurun_out/prof_lstm_l1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full1.log 2>&1; echo "ncu full lstm rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 106 -c 1 -o gpurun_out/prof_conv_d0 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1; echo "ncu full conv rc=$?"
ls -la gpurun_out
