#!/bin/bash
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'fps', round(d['ms_per_step'],2), 'ms/step; e2e', round(d['e2e']['value'],1), '; lstm TF/s', round(d['roofline']['achieved'] or 0,1), '; step TF/s', round(d['roofline']['whole_step_tflops'],1))"; }
timeout 600 python bench.py --precision bf16x3 --steps 5 --warmup 3 --no-cpu 2>gpurun_out/x3.err | show "C2 bf16x3:"
timeout 600 python bench.py --size 1024 --unroll 16 --batch 1 --steps 5 --warmup 3 --no-cpu 2>gpurun_out/c5i.err | show "C5-shape infer (1024^2,T=16,B=1):"
timeout 900 python bench.py --mode train --size 1024 --unroll 16 --batch 1 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/c5t.err | show "C5 train (1024^2,T=16,B=1/GPU):"
tail -3 gpurun_out/x3.err gpurun_out/c5i.err gpurun_out/c5t.err
nvidia-smi --query-gpu=memory.used --format=csv
