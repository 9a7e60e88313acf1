#!/bin/bash
# ncu --set full captures of every kernel class (function-name filters skip to the launch of interest)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 110 -c 1 -o gpurun_out/prof_lstm_l1 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu full lstm rc=$?"
timeout 900 $NCU -k regex:LuPpEdges -s 3 -c 1 -o gpurun_out/prof_pp_edges python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu pp edges rc=$?"
timeout 900 $NCU -k regex:LuPpFlattenBg -s 3 -c 1 -o gpurun_out/prof_pp_flatten python bench.py --mode postprocess --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu pp flatten rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 155 -c 1 -o gpurun_out/prof_conv_d0_c python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu conv rc=$?"
timeout 900 $NCU -k regex:LuUpsample2x -s 11 -c 1 -o gpurun_out/prof_upsample python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu upsample rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lu_conv_tc_kernel -s 307 -c 1 -o gpurun_out/prof_dgrad_l1 python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu dgrad rc=$?"
timeout 1500 $NCU -k regex:lu_wgrad_tc_kernel -s 69 -c 1 -o gpurun_out/prof_wgrad_l1 python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu wgrad rc=$?"
timeout 1500 $NCU -k regex:LuLstmCellBwd -s 95 -c 1 -o gpurun_out/prof_cellbwd python bench.py --mode train --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu cellbwd rc=$?"
ls -la gpurun_out/*.ncu-rep
