#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'fps', round(d['ms_per_step'],2), 'ms; lstm', round(d['roofline']['kernel_ms_per_step'],2), 'ms', round(d['roofline']['achieved'],1), 'TF/s; clocks', d['clocks']['sm_mhz'])"; done
