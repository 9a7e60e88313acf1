#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_wgrad_tc_kernel -s 94 -c 1 -o gpurun_out/prof_wgrad python bench.py --mode train --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_wg.log 2>&1; echo "ncu rc=$?"
