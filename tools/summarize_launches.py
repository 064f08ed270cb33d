#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per
kernel (count, total, mean, share of the profiled region) as a markdown table.

    python tools/summarize_launches.py gpurun_out/r1a_launches.csv "title" > profiles/r1a_launches.md
"""
import collections
import csv
import sys


def main(path, title=''):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    grid = {}
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1., 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
        k = row['Kernel Name'].split('(')[0].replace('void ', '')
        tot[k] += v
        cnt[k] += 1
        grid[k] = (row['Grid Size'], row['Block Size'])
    T = sum(tot.values())
    print('# ncu launch list: ' + title)
    print()
    print('source: `%s` (cold-cache, serialised launches: compare shares, not '
          'absolutes); %d launches, %.3f ms in total' % (path, sum(cnt.values()), T/1e3))
    print()
    print('| kernel | launches | total ms | mean us | share | last grid x block |')
    print('|---|---:|---:|---:|---:|---|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('| `%s` | %d | %.3f | %.2f | %.3f | %s x %s |' %
              (k, cnt[k], v/1e3, v/cnt[k], v/T, grid[k][0], grid[k][1]))


if __name__ == '__main__':
    main(*sys.argv[1:3])
