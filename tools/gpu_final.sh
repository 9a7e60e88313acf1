#!/bin/bash
# end-of-iteration evidence run: tests, smoke, every bench mode, launch lists (ncu --set full captures: tools/gpu_final2.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-160 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_infer.json
timeout 600 python bench.py --steps 10 --warmup 3 --post --no-cpu > gpurun_out/bench_infer_post.json 2>/dev/null; echo "bench post-e2e rc=$?"
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-160 gpurun_out/bench_train.json
timeout 300 python bench.py --mode stream --steps 300 --warmup 10 --no-cpu --post > gpurun_out/bench_stream.json 2>/dev/null; cut -c1-160 gpurun_out/bench_stream.json
timeout 300 python bench.py --mode postprocess --steps 400 --warmup 5 > gpurun_out/bench_post.json 2>/dev/null; cut -c1-160 gpurun_out/bench_post.json
timeout 300 python bench.py --mode augment --steps 100 --warmup 5 > gpurun_out/bench_aug.json 2>/dev/null; cut -c1-160 gpurun_out/bench_aug.json
timeout 300 python tools/time_metrics.py > gpurun_out/metrics_timing.json 2>/dev/null; cat gpurun_out/metrics_timing.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list train rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_post.csv python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list post rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_aug.csv python bench.py --mode augment --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list aug rc=$?"
