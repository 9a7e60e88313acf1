"""Copy the evidence of the last `tools/gpu_final.sh` run from gpurun_out/ (scratch) into profiles/ (tracked) and print
the per-kernel tables used in profiles/README.md."""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r1'

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__cluster_dim_x', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum']


def ncu_summary(rep, out):
    r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        return
    with open(out, 'w') as f:
        for h, u, v in zip(rows[0], rows[1], rows[2]):
            hh = h.split('TriageCompute.')[-1]
            if hh in KEEP:
                f.write('%-90s %-16s %s\n' % (hh, u, v))


def launch_table(path, step_index=4):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', ''))) for r in rows]
    starts = [i for i, (n, v) in enumerate(names) if 'LuPrepPatches' in n]
    fw = names[starts[step_index]:starts[step_index + 1]]
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


for src, dst in [('bench_infer.json', '%s_bench_n1_infer.json'), ('bench_train.json', '%s_bench_n1_train.json'),
                 ('bench_ref.json', '%s_bench_n1_reference_arm.json'), ('launches_infer.csv', '%s_launches_bench_steps1.csv'),
                 ('launches_train.csv', '%s_launches_train_steps1.csv'), ('bench_n2.json', '%s_bench_n2_infer.json'),
                 ('bench_train_n2.json', '%s_bench_n2_train.json'), ('bench_infer_post.json', '%s_bench_n1_infer_labelled.json'),
                 ('bench_stream.json', '%s_bench_n1_stream.json'), ('bench_post.json', '%s_bench_n1_postprocess.json'),
                 ('bench_aug.json', '%s_bench_n1_augment.json'), ('metrics_timing.json', '%s_metrics_timing.json'),
                 ('launches_post.csv', '%s_launches_postprocess_steps1.csv'), ('launches_aug.csv', '%s_launches_augment_steps1.csv')]:
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst % TAG))
for rep, out in [('prof_lstm_l1.ncu-rep', '%s_ncu_prof_lstm_l1.txt'), ('prof_wgrad_l1.ncu-rep', '%s_ncu_prof_wgrad_l1.txt'),
                 ('prof_conv_d0_c.ncu-rep', '%s_ncu_prof_conv_d0.txt'), ('prof_pp_edges.ncu-rep', '%s_ncu_prof_pp_edges.txt'),
                 ('prof_pp_flatten.ncu-rep', '%s_ncu_prof_pp_flatten_bg.txt'), ('prof_upsample.ncu-rep', '%s_ncu_prof_upsample.txt'),
                 ('prof_dgrad_l1.ncu-rep', '%s_ncu_prof_dgrad_l1.txt'), ('prof_cellbwd.ncu-rep', '%s_ncu_prof_lstm_cell_bwd.txt')]:
    if os.path.exists(os.path.join(G, rep)):
        ncu_summary(os.path.join(G, rep), os.path.join(P, out % TAG))
print('## inference step')
launch_table(os.path.join(G, 'launches_infer.csv'))
print('\n## train step')
launch_table(os.path.join(G, 'launches_train.csv'))


def tail_table(path, pattern, count):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', ''))) for r in rows if re.search(pattern, r['Kernel Name'])][-count:]
    print('| kernel | us |\n|---|---|')
    for n, v in names:
        m = re.search(r'(Lu\w+|lu_pp_\w+)', n)
        print('| `%s` | %.1f |' % (m.group(1) if m else n[:40], v / 1e3))
    print('| total | %.1f |' % (sum(v for _, v in names) / 1e3))


if os.path.exists(os.path.join(G, 'launches_post.csv')):
    print('\n## post-processing call (32 frames)')
    tail_table(os.path.join(G, 'launches_post.csv'), r'LuPp|lu_pp_', 19)
if os.path.exists(os.path.join(G, 'launches_aug.csv')):
    print('\n## augmentation of one sequence chunk (8 frames)')
    tail_table(os.path.join(G, 'launches_aug.csv'), r'LuAug', 5)
