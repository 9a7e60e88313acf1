#!/bin/bash
# Round 2, call F: final single-GPU evidence -- suite, default bench line, reference arm as the driver runs it, stream mode,
# launch lists and ncu --set full captures of the dominant kernels (copied into profiles/ by tools/r2_collect.py).
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/f_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/f_pytest.log
run() { tag=$1; shift; timeout -k 10 1500 "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err; echo "$tag rc=$? $(tail -1 gpurun_out/f_$tag.err)"; }
run bench python bench.py --steps 20 --warmup 5
run reference python bench.py --impl reference --steps 20 --warmup 5
run stream python bench.py --mode stream --no-parity --no-variants --steps 200 --warmup 20 --no-cpu
run train python bench.py --mode train --no-parity --steps 10 --warmup 3 --no-cpu
NCUL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout -k 10 900 $NCUL --log-file gpurun_out/f_launches_infer.csv python bench.py --mode infer --no-parity --no-variants --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "list infer rc=$?"
timeout -k 10 900 $NCUL --log-file gpurun_out/f_launches_train.csv python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "list train rc=$?"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
INF="python bench.py --mode infer --no-parity --no-variants --steps 1 --warmup 3 --no-cpu"
TRN="python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu"
cap() { tag=$1; shift; timeout -k 10 900 "$@" > /dev/null 2>&1; echo "ncu $tag rc=$?"; }
# forward of the 3rd warm-up call: per forward 32 ConvLSTM launches (<1,..>) in level order 8 + 8 + 8 + 8
cap lstm_l1 $NCU -k regex:"lu_conv_tc_kernel<1" -s 74 -c 1 -o gpurun_out/f_prof_lstm_l1 $INF
cap lstm_l3 $NCU -k regex:"lu_conv_tc_kernel<1" -s 90 -c 1 -o gpurun_out/f_prof_lstm_l3 $INF
cap conv_d0 $NCU -k regex:"lu_conv_tc_kernel<0" -s 42 -c 1 -o gpurun_out/f_prof_conv_d0 $INF
cap dgrad $NCU -k regex:"lu_conv_tc_kernel<2, true, 3" -s 40 -c 1 -o gpurun_out/f_prof_dgrad_pair $TRN
cap wgrad_pair $NCU -k regex:lu_wgrad_pair_kernel -s 30 -c 1 -o gpurun_out/f_prof_wgrad_pair $TRN
cap bnapply $NCU -k regex:LuBnApply -s 64 -c 1 -o gpurun_out/f_prof_bnapply $TRN
cap bnbwdreduce $NCU -k regex:LuBnBwdReduce -s 79 -c 1 -o gpurun_out/f_prof_bnbwdreduce $TRN
cap bnbwdapply $NCU -k regex:LuBnBwdApply -s 79 -c 1 -o gpurun_out/f_prof_bnbwdapply $TRN
cap cellbwd $NCU -k regex:LuLstmCellBwd -s 128 -c 1 -o gpurun_out/f_prof_cellbwd $TRN
ls -la gpurun_out/f_prof_*.ncu-rep 2>/dev/null | wc -l
python - <<'PY'
import json
for tag in ('bench','reference','stream','train'):
    try:
        d=json.loads(open('gpurun_out/f_%s.json'%tag).read()); t=d.get('train',{})
        print(tag, 'value %.2f ms %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'train %.1f fps %.1f ms' % (t.get('value',0), t.get('ms_per_step',0)), (d.get('roofline') or {}).get('frac'), d.get('clocks'))
    except Exception as e: print(tag, 'unreadable', e)
PY
