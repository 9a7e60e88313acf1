#!/bin/bash
# second-session validation: full GPU suite, smoke, post-processing bench + launch list, headline benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --mode postprocess --steps 20 --warmup 3 > gpurun_out/bench_post.json 2> gpurun_out/bench_post.err; echo "bench post rc=$?"; cut -c1-1200 gpurun_out/bench_post.json; tail -3 gpurun_out/bench_post.err
timeout 300 python bench.py --mode postprocess --steps 20 --warmup 3 --batch 1 --unroll 1 --no-cpu > gpurun_out/bench_post_b1.json 2>/dev/null; cut -c1-400 gpurun_out/bench_post_b1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_post.csv python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list post rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_infer.json
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-300 gpurun_out/bench_train.json
timeout 300 python bench.py --mode stream --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_stream.json 2>/dev/null; cut -c1-300 gpurun_out/bench_stream.json
