#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --mode postprocess --steps 400 --warmup 5 > gpurun_out/bench_post.json 2> gpurun_out/bench_post.err; echo "bench post rc=$?"; cut -c1-300 gpurun_out/bench_post.json; tail -3 gpurun_out/bench_post.err
timeout 300 python bench.py --mode postprocess --steps 2000 --warmup 5 --batch 1 --unroll 1 --no-cpu > gpurun_out/bench_post_b1.json 2>gpurun_out/bench_post_b1.err; cut -c1-300 gpurun_out/bench_post_b1.json; tail -3 gpurun_out/bench_post_b1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_post.csv python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list post rc=$?"
for g in off on; do timeout 300 python bench.py --mode stream --steps 300 --warmup 10 --no-cpu --cuda-graph $g > gpurun_out/bench_stream_$g.json 2>gpurun_out/bench_stream_$g.err; cut -c1-250 gpurun_out/bench_stream_$g.json; tail -2 gpurun_out/bench_stream_$g.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stream.csv python bench.py --mode stream --steps 1 --warmup 3 --no-cpu --cuda-graph off > /dev/null 2>&1; echo "ncu list stream rc=$?"
