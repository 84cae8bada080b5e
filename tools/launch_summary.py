"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py FILE [skip]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
d = collections.defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r'\(.*', '', r[ki])[:100]
    d[name][0] += 1
    d[name][1] += v
tot = sum(v[1] for v in d.values())
print('launches %d, total %.3f ms' % (sum(v[0] for v in d.values()), tot / 1e6))
for name, (c, t) in sorted(d.items(), key=lambda x: -x[1][1])[:30]:
    print('%9.3f ms %5.1f%% %6d x %8.1f us  %s' % (t / 1e6, 100 * t / tot, c, t / c / 1e3, name))
