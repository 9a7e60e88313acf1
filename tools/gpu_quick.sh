#!/bin/bash
# quick iteration check: GPU suite, headline benches with / without the wide cluster multicast
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 1 0; do
LU_CLUSTER_WIDE=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_infer_w$w.json 2> gpurun_out/bench_infer_w$w.err; echo "bench infer wide=$w rc=$?"; cut -c1-160 gpurun_out/bench_infer_w$w.json; tail -2 gpurun_out/bench_infer_w$w.err
LU_CLUSTER_WIDE=$w timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train_w$w.json 2> gpurun_out/bench_train_w$w.err; echo "bench train wide=$w rc=$?"; cut -c1-160 gpurun_out/bench_train_w$w.json; tail -2 gpurun_out/bench_train_w$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list train rc=$?"
