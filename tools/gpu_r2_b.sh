#!/bin/bash
# Round 2, call B: GPU suite (no -x: every failure listed), then the rest of the cta_group::2 bring-up.
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/b_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -8 gpurun_out/b_pytest.log
timeout -k 10 1800 bash tools/gpu_pair.sh > gpurun_out/b_pair.log 2>&1; echo "pair rc=$?"; grep "==\|STOP\|ALL STEPS\|rc=" gpurun_out/b_pair.log | tail -30
for f in gpurun_out/bench_*pair*.json gpurun_out/bench_train_wg*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(' value %.2f ms %.2f' % (d['value'], d['ms_per_step']), {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],2)) for k,v in d.get('train',{}).get('rooflines',{}).items()} or (round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms_per_step'],2)))
except Exception as e: print(' unreadable', e)
PY
done
