"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import collections
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
ui = hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    n += 1
    if n <= skip:
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    v_us = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    name = re.sub(r'\(.*', '', r[ki])
    agg[name][0] += 1
    agg[name][1] += v_us
    tot += v_us
print('total %.1f us over %d launches' % (tot, n - skip))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%8.1f us  %5.1f%%  x%-4d  %s' % (t, 100 * t / tot, c, k[:100]))
