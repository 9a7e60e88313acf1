#!/bin/bash
# Round 2, call M: per-column constants staged once for single-N-tile convolutions: parity, then same-box A/B (LU_CST_PER_TILE).
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ctc_parity.py tests/test_gpu_blocks.py -m gpu -q -x 2>&1 | tail -3
i=0
for cfg in "LU_CST_PER_TILE=1" "LU_CST_PER_TILE=0" "LU_CST_PER_TILE=1" "LU_CST_PER_TILE=0"; do
  i=$((i+1))
  env $cfg timeout -k 10 300 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu > gpurun_out/m_$i.json 2> gpurun_out/m_$i.err
  python - "$cfg" gpurun_out/m_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print('%-20s'%sys.argv[1], 'infer %.2f fps %.2f ms | lstm %.2f ms | e2e %.1f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value']), d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'unreadable',e)
PY
done
for cfg in "LU_CST_PER_TILE=1" "LU_CST_PER_TILE=0"; do
  i=$((i+1))
  env $cfg timeout -k 10 300 python bench.py --mode train --no-parity --steps 6 --warmup 3 --no-cpu > gpurun_out/m_$i.json 2> gpurun_out/m_$i.err
  python - "$cfg" gpurun_out/m_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read()); t=d['train']
    print('%-20s'%sys.argv[1], 'train %.1f ms'%t['ms_per_step'], {k:(round(v['kernel_ms_per_step'],1)) for k,v in t['rooflines'].items()}, 'elem %.1f'%t['elementwise_and_other_ms_per_step'], d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1],'unreadable',e)
PY
done
