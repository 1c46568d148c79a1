#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel + SASS opcode mix (used to write profiles/*.md)."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum']


def main(path, kidx=0):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')][:110])
        for k in KEYS:
            if k in hdr:
                print('  %-72s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        for i, k in enumerate(hdr):
            if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                v = float(r[i])
                if v > 0.08:
                    print('  stall %-62s %.3f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = None
    data = []
    k = -1
    for r in rows:
        if r and r[0] == 'Kernel Name':
            k += 1
            continue
        if r and r[0] == 'Address':
            hdr = r
            continue
        if k == kidx:
            data.append(r)
    ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[ie]) for r in data)
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia])
        op = m.group(2) if m else '?'
        op = '.'.join(op.split('.')[:2]) if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'F2F', 'SHFL')) else op.split('.')[0]
        ops[op] += int(r[ie])
        samp[op] += int(r[isamp])
    print('total warp instructions', tot, ' SASS lines', len(data))
    for op, c in ops.most_common(28):
        print('  %-14s %12d %5.1f%%  stall-samples %d' % (op, c, 100.0 * c / tot, samp[op]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
