#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text report:  python ncu_summary.py rep [kernel-substr]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, 'raw')
    h, u = raw[0], raw[1]
    keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed']
    for row in raw[2:]:
        print('== kernel:', row[h.index('Kernel Name')][:110])
        for k in keys:
            if k in h:
                print('   %-70s %-14s %s' % (k, u[h.index(k)], row[h.index(k)]))
    src = page(rep, 'source')
    hdr = None
    for i, r in enumerate(src):
        if r and r[0] == 'Address':
            hdr, data = r, src[i + 1:]
            break
    if hdr is None:
        return
    ix = {c: i for i, c in enumerate(hdr)}
    byop, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    scols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
    for r in data:
        if len(r) < len(hdr):
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
        op = (m.group(2) if m else r[ix['Source']][:10]).split('.')[0]
        ne = int(r[ix['Instructions Executed']] or 0)
        tot += ne
        byop[op] += ne
        samp[op] += int(r[ix['# Samples']] or 0)
        for c in scols:
            if r[ix[c]]:
                stall[c] += int(r[ix[c]])
    print('== SASS mix (warp instructions, first kernel): total %d' % tot)
    for k, v in byop.most_common(16):
        print('   %-10s %14d %5.1f%%   samples %d' % (k, v, 100.0 * v / tot, samp[k]))
    ts = sum(stall.values())
    print('== warp stall samples')
    for k, v in stall.most_common(9):
        print('   %-26s %9d %5.1f%%' % (k, v, 100.0 * v / ts))


if __name__ == '__main__':
    main()
