#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel and the most-stalled SASS instructions.
usage: tools/ncu_summary.py report.ncu-rep [kernel-regex] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if name in seen:
        continue
    seen.add(name)
    print('== ' + name)
    for w in WANT:
        if w in hdr:
            print('   %-80s %s %s' % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv']
if kre:
    cmd += ['--kernel-name', 'regex:' + kre]
src = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
blocks = []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r
        blocks.append([])
    elif hdr and r and r[0].startswith('0x') and len(r) >= len(hdr) - 2:
        blocks[-1].append(r)
if blocks:
    data = blocks[0]
    col = {n: hdr.index(n) for n in ('# Samples', 'Source', 'stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_mio',
                                       'stall_lg', 'stall_math', 'stall_wait', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal',
                                       'L2 Theoretical Sectors Global', 'L2 Theoretical Sectors Global Ideal')}
    tot = sum(int(r[col['# Samples']]) for r in data)
    print('-- first launch of the source page: %d samples over %d SASS instructions' % (tot, len(data)))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][col['# Samples']]))[:topn]
    for i in sorted(idx):
        r = data[i]
        print('%5d %-52s %5.1f%% long %s bar %s short %s mio %s lg %s math %s | shW %s/%s gl %s/%s' % (
            i, r[col['Source']].strip()[:52], 100.0 * int(r[col['# Samples']]) / max(tot, 1), r[col['stall_long_sb']],
            r[col['stall_barrier']], r[col['stall_short_sb']], r[col['stall_mio']], r[col['stall_lg']], r[col['stall_math']],
            r[col['L1 Wavefronts Shared']], r[col['L1 Wavefronts Shared Ideal']], r[col['L2 Theoretical Sectors Global']],
            r[col['L2 Theoretical Sectors Global Ideal']]))
