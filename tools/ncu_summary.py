"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/launches_rNN.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/kernel_rNN.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']


def launches(path):
    rows = list(csv.reader(open(path, errors='ignore')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[ci['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        if v != v:          # 'nan': a launch ncu could not time
            continue
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ci['Metric Unit']], 1.0)
        name = r[ci['Kernel Name']][:90]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    print('# ncu --metrics gpu__time_duration.sum launch list: %d launches, %.2f ms of kernel time' % (n, tot))
    print('# (cold-cache, serialised: compare SHARES, not absolutes)')
    print('%10s %6s %7s  %s' % ('ms', 'share', 'count', 'kernel'))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print('%10.3f %5.1f%% %7d  %s' % (v, 100 * v / tot, c, k))


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('# ncu --set full --clock-control none : %s' % path)
        for k in KEYS:
            if k in ci:
                print('%-78s %s %s' % (k, r[ci[k]], units[ci[k]]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
