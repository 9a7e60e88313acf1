#!/bin/bash
# Round 2, call I: weight-gradient pair kernel: 16-byte reductions in the flush (and LU_WGRAD_PAIR=3 tap-pair entries): parity
# against the scalar engine, then A/B bench of the train step.
mkdir -p gpurun_out
for m in 1 3; do LU_WGRAD_PAIR=$m timeout -k 10 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_ctc_parity.py -m gpu -q -x -k "wgrad or weight_gradient" 2>&1 | tail -2; done
for m in 1 3; do LU_WGRAD_PAIR=$m timeout -k 10 300 python bench.py --mode train --no-parity --steps 6 --warmup 3 --no-cpu > gpurun_out/i_train_pair$m.json 2> gpurun_out/i_train_pair$m.err; echo "pair$m rc=$?"; done
python - <<'PY'
import json
for m in (1,3):
    try:
        d=json.loads(open('gpurun_out/i_train_pair%d.json'%m).read()); t=d['train']
        print('LU_WGRAD_PAIR=%d'%m, '%.1f ms'%t['ms_per_step'], {k:(round(v['kernel_ms_per_step'],1), round(v['frac'],3)) for k,v in t['rooflines'].items()}, d['config'].get('switches'))
    except Exception as e: print(m,'unreadable',e)
PY
