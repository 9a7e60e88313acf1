#!/bin/bash
# Round 2, call E: the CTA-pair weight-gradient kernel (transposed product): suite, A/B of LU_WGRAD_PAIR, ncu capture.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_ctc_parity.py -m gpu -q -x -k "wgrad or weight_gradient or train_step" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/e_pytest_wgrad.log; rc=${PIPESTATUS[0]}; echo "pytest wgrad rc=$rc"; tail -6 gpurun_out/e_pytest_wgrad.log
if [ $rc -ne 0 ]; then
  for nb in 1 0; do LU_WGRAD_PAIR=$nb timeout -k 10 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "wgrad" 2>&1 | tail -3; done
  echo "STOP: pair wgrad kernel fails"; exit 1
fi
timeout -k 10 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/e_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/e_pytest.log
run() { tag=$1; shift; timeout -k 10 600 "$@" > gpurun_out/e_$tag.json 2> gpurun_out/e_$tag.err; echo "$tag rc=$?"; }
for nb in 0 1 2; do LU_WGRAD_PAIR=$nb run train_wgpair$nb python bench.py --mode train --no-parity --steps 8 --warmup 3 --no-cpu; done
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:lu_wgrad_pair_kernel -s 36 -c 1 -o gpurun_out/e_prof_wgrad_pair python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu wgrad pair rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_train.csv python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
for f in gpurun_out/e_train_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    t=d.get('train',{})
    print(' value %.2f ms %.2f' % (d['value'], d['ms_per_step']), 'elem %.2f' % t.get('elementwise_and_other_ms_per_step',0), {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],2)) for k,v in t.get('rooflines',{}).items()})
except Exception as e: print(' unreadable', e)
PY
done
