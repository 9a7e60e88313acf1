#!/bin/bash
# Round 2, call J: A/B on ONE box of the weight-gradient flush (LU_WGRAD_RED4) and of the pixel range per task (LU_WGRAD_RANGE).
mkdir -p gpurun_out
i=0
for cfg in "LU_WGRAD_RED4=0" "LU_WGRAD_RED4=1" "LU_WGRAD_RANGE=512" "LU_WGRAD_RANGE=1024" "LU_WGRAD_RED4=0" "LU_WGRAD_RED4=1"; do
  i=$((i+1))
  env $cfg timeout -k 10 300 python bench.py --mode train --no-parity --steps 6 --warmup 3 --no-cpu > gpurun_out/j_$i.json 2> gpurun_out/j_$i.err
  python - "$cfg" gpurun_out/j_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read()); t=d['train']
    print('%-22s'%sys.argv[1], '%.1f ms'%t['ms_per_step'], {k:(round(v['kernel_ms_per_step'],1), round(v['frac'],3)) for k,v in t['rooflines'].items()}, 'elem %.1f'%t['elementwise_and_other_ms_per_step'], d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'unreadable',e)
PY
done
