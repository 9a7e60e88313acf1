#!/bin/bash
# quick iteration check: GPU suite, headline bench, post-processing bench + launch lists, two remaining ncu captures
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_infer.json
timeout 300 python bench.py --mode postprocess --steps 400 --warmup 5 > gpurun_out/bench_post.json 2>/dev/null; cut -c1-160 gpurun_out/bench_post.json
timeout 300 python tools/time_metrics.py > gpurun_out/metrics_timing.json 2>/dev/null; cut -c1-200 gpurun_out/metrics_timing.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_post.csv python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list post rc=$?"
timeout 900 $NCU -k regex:lu_conv_tc_kernel -s 155 -c 1 -o gpurun_out/prof_conv_d0_c python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu conv rc=$?"
timeout 1500 $NCU -k regex:lu_conv_tc_kernel -s 307 -c 1 -o gpurun_out/prof_dgrad_l1 python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu dgrad rc=$?"
