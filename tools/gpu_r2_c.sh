#!/bin/bash
# Round 2, call C: GPU suite on the fresh library (pair mode default, row-loop element-wise kernels), train A/B of LU_PAIR,
# default bench line, launch list of one train step.
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/c_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/c_pytest.log
for p in 0 1; do
  LU_PAIR=$p timeout -k 10 600 python bench.py --mode train --no-parity --steps 8 --warmup 3 --no-cpu > gpurun_out/c_train_pair$p.json 2> gpurun_out/c_train_pair$p.err; echo "train pair=$p rc=$?"
done
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c_bench.err
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_train.csv python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
for f in gpurun_out/c_train_pair0.json gpurun_out/c_train_pair1.json gpurun_out/c_bench.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    t=d.get('train',{})
    print(' value %.2f ms %.2f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'train ms %.2f elem %.2f' % (t.get('ms_per_step',0), t.get('elementwise_and_other_ms_per_step',0)), {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],2)) for k,v in t.get('rooflines',{}).items()}, round(d['roofline']['frac'],3), d.get('clocks'))
except Exception as e: print(' unreadable', e)
PY
done
