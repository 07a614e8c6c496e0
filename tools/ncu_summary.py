#!/usr/bin/env python
"""Summarise an .ncu-rep: per-launch headline metrics (raw page) and, for one kernel, the SASS opcode mix and stall reasons (source page)."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[h.index('Kernel Name')], 'id', r[0])
        for k in KEYS:
            if k in h:
                print('   %-70s %s %s' % (k, r[h.index(k)], units[h.index(k)]))


def source(rep, kernel, skip=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel, '--launch-skip', str(skip),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][:2])
    h = rows[1]
    data = [r for r in rows[2:] if len(r) == len(h)]
    si, ei, sm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    stall = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    tot = sum(int(r[ei]) for r in data if r[ei].isdigit())
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        if not r[ei].isdigit():
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
        op = m.group(2).split('.')[0] if m else '?'
        ops[op] += int(r[ei])
        samp[op] += int(r[sm]) if r[sm].isdigit() else 0
    print('total warp instructions', tot, 'static', len(data))
    for op, c in ops.most_common(28):
        print('  %-10s %6.2f%%  samples %d' % (op, 100.0 * c / tot, samp[op]))
    st = collections.Counter()
    for r in data:
        for i in stall:
            if r[i].isdigit():
                st[h[i]] += int(r[i])
    ts = sum(st.values()) or 1
    print({k: round(100.0 * v / ts, 1) for k, v in st.most_common(10)})
    return data, h


if __name__ == '__main__':
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        source(rep, sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        raw(rep)
