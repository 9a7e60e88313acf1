"""Copy the round-2 evidence of tools/gpu_r2_f.sh / gpu_r2_n2.sh / gpu_r2_n8.sh from gpurun_out/ (scratch) into profiles/
(tracked) and print the per-kernel tables used in profiles/README.md.     python tools/r2_collect.py"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ncu_summary      # noqa: E402

COPY = [('f_bench.json', 'r2_bench_n1_default.json'), ('f_reference.json', 'r2_bench_n1_reference_arm.json'),
        ('f_stream.json', 'r2_bench_n1_stream.json'), ('f_train.json', 'r2_bench_n1_train.json'),
        ('f_launches_infer.csv', 'r2_launches_infer_steps1.csv'), ('f_launches_train.csv', 'r2_launches_train_steps1.csv'),
        ('n2_bench.json', 'r2_bench_n2_default.json'), ('n2_train_syncbn.json', 'r2_bench_n2_train_syncbn.json'),
        ('n8_bench.json', 'r2_bench_n8_default.json'), ('n8_c5_train.json', 'r2_bench_n8_c5_train.json'),
        ('n8_c5_infer.json', 'r2_bench_n8_c5_infer.json'), ('n8_train_syncbn.json', 'r2_bench_n8_train_syncbn.json')]
NCU = [('f_prof_lstm_l1', 'r2_ncu_prof_lstm_l1_pair'), ('f_prof_lstm_l3', 'r2_ncu_prof_lstm_l3_pair'),
       ('f_prof_conv_d0', 'r2_ncu_prof_conv_d0'), ('f_prof_dgrad_pair', 'r2_ncu_prof_dgrad_pair'),
       ('f_prof_wgrad_pair', 'r2_ncu_prof_wgrad_pair'), ('f_prof_bnapply', 'r2_ncu_prof_bn_apply'),
       ('f_prof_bnbwdreduce', 'r2_ncu_prof_bn_bwd_reduce'), ('f_prof_bnbwdapply', 'r2_ncu_prof_bn_bwd_apply'),
       ('f_prof_cellbwd', 'r2_ncu_prof_lstm_cell_bwd')]


def launch_table(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', ''))) for r in rows]
    starts = [i for i, (n, v) in enumerate(names) if 'LuPrepPatches' in n]
    fw = names[starts[-2]:starts[-1]]          # one whole step period (weight re-packing of the step included)
    tot = collections.OrderedDict()
    for n, v in fw:
        key = re.sub(r'\(.*', '', n).replace('void ', '')[:60]
        tot.setdefault(key, [0, 0.0])
        tot[key][0] += 1
        tot[key][1] += v
    s = sum(v for _, v in tot.values())
    print('| kernel | launches | ms | share |\n|---|---|---|---|')
    for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if v / s > 0.002:
            print('| `%s` | %d | %.2f | %.1f %% |' % (k, c, v / 1e6, 100 * v / s))
    print('| total | %d | %.1f | |' % (len(fw), s / 1e6))


for src, dst in COPY:
    if os.path.exists(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
        print('copied', dst)
for rep, out in NCU:
    if os.path.exists(os.path.join(G, rep + '.ncu-rep')):
        ncu_summary.main(os.path.join(G, rep + '.ncu-rep'), os.path.join(P, out + '.txt'))
        print('summarised', out)
for tag in ('infer', 'train'):
    p = os.path.join(G, 'f_launches_%s.csv' % tag)
    if os.path.exists(p):
        print('\n## one %s step' % tag)
        launch_table(p)
for f in sorted(os.listdir(P)):
    if f.startswith('r2_bench') and f.endswith('.json'):
        try:
            d = json.loads(open(os.path.join(P, f)).read())
            t = d.get('train') or {}
            print('%-44s value %9.2f %s | e2e %9.2f | train %7.2f fps %7.2f ms' % (f, d['value'], d['unit'], d['e2e']['value'],
                                                                                 t.get('value', 0), t.get('ms_per_step', 0)))
        except Exception as e:
            print(f, 'unreadable', e)
