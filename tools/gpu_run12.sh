#!/bin/bash
timeout 300 python bench.py --mode stream --steps 64 --warmup 8 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('stream B=1 T=1:', round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms/frame; e2e', round(d['e2e']['value'],1), 'fps', round(d['e2e']['ms_per_step'],3),'ms; launches/step', d['gpu_launches']/d['steps'])"
