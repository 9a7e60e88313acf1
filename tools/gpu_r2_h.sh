#!/bin/bash
# Round 2, call H: the bench lines of call F (whose `/usr/bin/time` wrapper does not exist on the box) and the four
# lu_conv_tc_kernel captures, selected by launch index among ALL lu_conv_tc_kernel launches (indices read from
# gpurun_out/f_launches_{infer,train}.csv: 49 launches per inference forward, 109 per train step).
mkdir -p gpurun_out
run() { tag=$1; shift; t0=$SECONDS; timeout -k 10 1500 "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err; echo "$tag rc=$? wall $((SECONDS-t0)) s"; echo "wall_s $((SECONDS-t0))" >> gpurun_out/f_$tag.err; }
run bench python bench.py --steps 20 --warmup 5
run reference python bench.py --impl reference --steps 20 --warmup 5
run stream python bench.py --mode stream --no-parity --no-variants --steps 200 --warmup 20 --no-cpu
run train python bench.py --mode train --no-parity --steps 10 --warmup 3 --no-cpu
NCU="ncu --set full --clock-control none --import-source on"
INF="python bench.py --mode infer --no-parity --no-variants --steps 1 --warmup 3 --no-cpu"
TRN="python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu"
cap() { tag=$1; shift; timeout -k 10 900 "$@" > gpurun_out/cap_$tag.log 2>&1; echo "ncu $tag rc=$? $(grep -c 'PROF' gpurun_out/cap_$tag.log) prof lines"; }
cap lstm_l1 $NCU -k regex:lu_conv_tc_kernel -s 110 -c 1 -o gpurun_out/f_prof_lstm_l1 $INF
cap lstm_l3 $NCU -k regex:lu_conv_tc_kernel -s 130 -c 1 -o gpurun_out/f_prof_lstm_l3 $INF
cap conv_d0 $NCU -k regex:lu_conv_tc_kernel -s 106 -c 1 -o gpurun_out/f_prof_conv_d0 $INF
cap dgrad $NCU -k regex:lu_conv_tc_kernel -s 307 -c 1 -o gpurun_out/f_prof_dgrad_pair $TRN
ls -la gpurun_out/f_prof_*.ncu-rep 2>/dev/null | wc -l
python - <<'PY'
import json
for tag in ('bench','reference','stream','train'):
    try:
        d=json.loads(open('gpurun_out/f_%s.json'%tag).read()); t=d.get('train',{})
        print(tag, 'value %.2f ms %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'train %.1f fps %.1f ms' % (t.get('value',0), t.get('ms_per_step',0)), (d.get('roofline') or {}).get('frac'), d.get('clocks'))
    except Exception as e: print(tag, 'unreadable', e)
PY
