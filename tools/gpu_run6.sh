#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 106 -c 1 -o gpurun_out/prof_conv_d0_c python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1; echo "ncu full conv rc=$?"
