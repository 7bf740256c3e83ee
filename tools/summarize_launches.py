"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    unit = r[mu]
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3 if unit in ('ms', 'msecond') else v)
    name = r[kn].split('(')[0]
    agg[name][0] += 1; agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f'total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches')
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{us/1e3:9.3f} ms {100*us/tot:5.1f}%  n={n:5d}  avg {us/n:8.1f} us  {name[:90]}')
