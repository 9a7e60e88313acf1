#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_tc.py two tcgen05 halo bf16x3 2 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cl in 2 1 2 1; do
LU_CLUSTER=$cl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cluster=$cl', round(d['value'],1), 'fps', round(d['ms_per_step'],2), 'ms; lstm', round(d['roofline']['kernel_ms_per_step'],2), 'ms', round(d['roofline']['achieved'],1), 'TF/s; clocks', d['clocks']['sm_mhz'])"
done
