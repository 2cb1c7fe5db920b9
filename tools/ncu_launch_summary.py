#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per
kernel name.  python tools/ncu_launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        n += 1
        name = row['Kernel Name'][:72]
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        if unit in ('nsecond', 'ns'):
            v /= 1000
        elif unit in ('msecond', 'ms'):
            v *= 1000
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'{n} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k:72s} n={len(v):4d} avg={sum(v) / len(v):9.1f}us tot={sum(v):10.1f}us '
              f'{100 * sum(v) / tot:5.1f}%')


if __name__ == '__main__':
    main()
