"""Kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: python tools/launch_shares.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0].replace('void ', '')
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(',', ''))
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f'{n:58s} {c:4d} launches {t / 1e3:9.1f} us {100 * t / tot:5.1f} %  avg {t / c / 1e3:8.1f} us')
print(f'total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches')
