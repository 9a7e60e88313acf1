#!/bin/bash
# end-of-iteration evidence run: tests, smoke, both benches, launch lists, one ncu --set full of the ConvLSTM kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_infer.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-300 gpurun_out/bench_train.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list train rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 110 -c 1 -o gpurun_out/prof_lstm_l1 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu full lstm rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_wgrad_tc_kernel -s 94 -c 1 -o gpurun_out/prof_wgrad_l1 python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu full wgrad rc=$?"
