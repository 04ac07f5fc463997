"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch table + per-kernel totals."""
import collections
import csv
import sys

path = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 80
lines = [l for l in open(path) if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit in ('ns', 'nsecond'):
        v /= 1e3
    elif unit in ('ms', 'msecond'):
        v *= 1e3
    seq.append((name, v, row['Grid Size']))
print(len(seq), 'launches')
for s in seq[first:first + count]:
    print('%-42s %10.1f us  %s' % (s[0][:42], s[1], s[2]))
agg = collections.OrderedDict()
for n, v, _ in seq:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print('--- totals over the capture ---')
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-42s n=%4d  %10.1f us  %5.1f%%' % (n[:42], c, v, 100 * v / tot))
