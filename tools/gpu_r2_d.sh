#!/bin/bash
# Round 2, call D: suite after the cell-backward fix + multi-channel path, threshold experiments for the clustered conv
# launches, train launch list, ncu of two reduction kernels.
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/d_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/d_pytest.log
run() { tag=$1; shift; timeout -k 10 600 "$@" > gpurun_out/d_$tag.json 2> gpurun_out/d_$tag.err; echo "$tag rc=$?"; }
run infer_base python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu
LU_WIDE_MIN_K=1024 run infer_k1024 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu
LU_WIDE_MIN_K=1024 LU_WIDE_MIN_BN=64 run infer_k1024_bn64 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu
LU_WIDE_MIN_K=512 LU_WIDE_MIN_BN=32 run infer_k512_bn32 python bench.py --mode infer --no-parity --no-variants --steps 10 --warmup 3 --no-cpu
run train_base python bench.py --mode train --no-parity --steps 8 --warmup 3 --no-cpu
LU_WIDE_MIN_K=1024 LU_WIDE_MIN_BN=64 run train_k1024_bn64 python bench.py --mode train --no-parity --steps 8 --warmup 3 --no-cpu
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_train.csv python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:LuBnStats -s 64 -c 1 -o gpurun_out/d_prof_bnstats python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu bnstats rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:LuBnBwdReduce -s 79 -c 1 -o gpurun_out/d_prof_bnbwdreduce python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu bnbwdreduce rc=$?"
for f in gpurun_out/d_infer_*.json gpurun_out/d_train_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    t=d.get('train',{})
    print(' value %.2f ms %.2f' % (d['value'], d['ms_per_step']), 'elem %.2f' % t.get('elementwise_and_other_ms_per_step',0), {k:(round(v['frac'],3), round(v['kernel_ms_per_step'],2)) for k,v in t.get('rooflines',{}).items()}, round(d['roofline']['frac'],3), d['roofline']['kernel_ms_per_step'])
except Exception as e: print(' unreadable', e)
PY
done
