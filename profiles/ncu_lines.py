#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of an .ncu-rep captured with --import-source on:
    python ncu_lines.py rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == 'File Name':
            cur_file = r[1].split('/')[-1]
        elif r and r[0] == 'Line No':
            hdr = {c: i for i, c in enumerate(r)}
        elif hdr and r and r[0].isdigit():
            try:
                ne, ns = int(r[hdr['Instructions Executed']]), int(r[hdr['# Samples']])
            except ValueError:
                continue
            key = (cur_file, int(r[0]))
            e = lines.setdefault(key, [0, 0, r[1].strip()[:90]])
            e[0] += ne
            e[1] += ns
    tot_i = sum(v[0] for v in lines.values())
    tot_s = sum(v[1] for v in lines.values())
    print('total warp instructions %d, samples %d' % (tot_i, tot_s))
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * v[1] / max(tot_s, 1), 100.0 * v[0] / max(tot_i, 1), f, ln, v[2]))


if __name__ == '__main__':
    main()
