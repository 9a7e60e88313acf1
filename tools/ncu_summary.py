"""ncu report (.ncu-rep, `ncu --set full`) -> the short text summary kept under profiles/ (same metric list as round 1).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_ncu_x.txt"""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__cluster_dim_x', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'gpc__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second']


def main(rep, out):
    r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        raise SystemExit('no kernel in %s' % rep)
    with open(out, 'w') as f:
        for k in range(2, len(rows)):
            for h, u, v in zip(rows[0], rows[1], rows[k]):
                hh = h.split('TriageCompute.')[-1]
                if hh in KEEP:
                    f.write('%-90s %-16s %s\n' % (hh, u, v))
            f.write('\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
