#!/bin/bash
# Round 2, call Q: second MMA-issuing thread for the resident-weight (narrow) convolutions (LU_TWO_ISSUERS): parity, same-box A/B.
mkdir -p gpurun_out
LU_TWO_ISSUERS=1 timeout -k 10 200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_blocks.py tests/test_gpu_ctc_parity.py -m gpu -q -x 2>&1 | tail -3
i=0
for cfg in "LU_TWO_ISSUERS=0" "LU_TWO_ISSUERS=1" "LU_TWO_ISSUERS=0" "LU_TWO_ISSUERS=1"; do
  i=$((i+1))
  env $cfg timeout -k 10 200 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu > gpurun_out/q_$i.json 2> gpurun_out/q_$i.err
  python - "$cfg" gpurun_out/q_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print('%-20s'%sys.argv[1], 'infer %.2f fps %.2f ms | lstm %.2f ms -> rest %.2f ms'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step']), d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'unreadable',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
done
