#!/bin/bash
# quick iteration check: GPU suite, headline benches with / without the resident weight panel of the narrow convs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 1 0; do
LU_B_RESIDENT=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_infer_r$w.json 2> gpurun_out/bench_infer_r$w.err; echo "bench infer resident=$w rc=$?"; cut -c1-160 gpurun_out/bench_infer_r$w.json; tail -2 gpurun_out/bench_infer_r$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
LU_B_RESIDENT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer_r0.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list r0 rc=$?"
