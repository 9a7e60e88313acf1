#!/bin/bash
# Round 2, call K: whole GPU suite + the bench lines + launch lists on the build with the 16-byte weight-gradient flush.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -6 > gpurun_out/k_pytest.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/k_pytest.log | cut -c1-300
run() { tag=$1; shift; t0=$SECONDS; timeout -k 10 900 "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err; echo "$tag rc=$? wall $((SECONDS-t0)) s"; }
run bench python bench.py --steps 20 --warmup 5
run train python bench.py --mode train --no-parity --steps 10 --warmup 3 --no-cpu
run stream python bench.py --mode stream --no-parity --no-variants --steps 200 --warmup 20 --no-cpu
NCUL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout -k 10 600 $NCUL --log-file gpurun_out/f_launches_infer.csv python bench.py --mode infer --no-parity --no-variants --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "list infer rc=$?"
timeout -k 10 600 $NCUL --log-file gpurun_out/f_launches_train.csv python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "list train rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lu_wgrad_pair_kernel -s 30 -c 1 -o gpurun_out/f_prof_wgrad_pair python bench.py --mode train --no-parity --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu wgrad rc=$?"
python - <<'PY'
import json
for tag in ('bench','train','stream'):
    try:
        d=json.loads(open('gpurun_out/f_%s.json'%tag).read()); t=d.get('train',{})
        print(tag, 'value %.2f ms %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'train %.1f fps %.1f ms' % (t.get('value',0), t.get('ms_per_step',0)), (d.get('roofline') or {}).get('frac'), d.get('clocks'))
    except Exception as e: print(tag, 'unreadable', e)
PY
