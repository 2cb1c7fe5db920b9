#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source sass,cuda --csv` output per CUDA
source line: warp-stall samples and executed instructions.

  python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    extra = sys.argv[4:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name',
                          f'regex:{kern}', '--print-source', 'sass,cuda'] + extra,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, agg, src = None, None, {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            i_s = hdr.index('Warp Stall Sampling (All Samples)')
            i_i = hdr.index('Instructions Executed')
            continue
        if hdr is None or len(r) <= i_i:
            continue
        try:
            line = int(r[0])
            s, n = int(r[i_s] or 0), int(r[i_i] or 0)
        except ValueError:
            continue
        k = (cur, line)
        a = agg.setdefault(k, [0, 0])
        a[0] += s
        a[1] += n
        src[k] = r[1]
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print(f'total stall samples {tot}, warp instructions {toti}')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'{a[0]:7d} {100 * a[0] / tot:5.1f}%  inst {100 * a[1] / toti:5.1f}%  {k[0]}:{k[1]:<4d} {src[k].strip()[:90]}')


if __name__ == '__main__':
    main()
