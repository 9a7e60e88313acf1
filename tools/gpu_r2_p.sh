#!/bin/bash
# Round 2, call P: CTA-pair mode (one M = 256 MMA per pair: half the issuer instructions per pixel) for the narrow convolutions
# too (LU_WIDE_MIN_K / LU_WIDE_MIN_BN): parity, then same-box A/B.
mkdir -p gpurun_out
LU_WIDE_MIN_K=256 LU_WIDE_MIN_BN=32 timeout -k 10 300 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ctc_parity.py -m gpu -q -x -k "not train" 2>&1 | tail -3
i=0
for cfg in "LU_WIDE_MIN_K=1024 LU_WIDE_MIN_BN=128" "LU_WIDE_MIN_K=256 LU_WIDE_MIN_BN=32" "LU_WIDE_MIN_K=256 LU_WIDE_MIN_BN=64" "LU_WIDE_MIN_K=1024 LU_WIDE_MIN_BN=128" "LU_WIDE_MIN_K=256 LU_WIDE_MIN_BN=32"; do
  i=$((i+1))
  env $cfg timeout -k 10 300 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu > gpurun_out/p_$i.json 2> gpurun_out/p_$i.err
  python - "$cfg" gpurun_out/p_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print('%-40s'%sys.argv[1], 'infer %.2f fps %.2f ms | lstm %.2f ms -> rest %.2f ms'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step']), d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'unreadable',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
done
