"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv` launch list: per kernel name the launch count,
total / average duration, share of the total, DRAM bytes per launch and tensor-pipe activity.
    python tools/launch_summary.py launches.csv [skip_first_n_launches] [traffic.json]
With a third argument, dram bytes (read + write) per launch and kernel are also written as JSON (bench.py: roofline.traffic)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
iid, kn, mn, mu, mv = H.index('ID'), H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Unit'), H.index('Metric Value')
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    d = launch.setdefault(int(r[iid]), {'name': r[kn]})
    unit = r[mu]
    if r[mn].startswith('gpu__time_duration'):
        v = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v * 1e6 if unit in ('s', 'second') else v
        d['us'] = v
    elif 'bytes_read' in r[mn]:
        d['rd'] = v * (1e3 if unit == 'Kbyte' else 1e6 if unit == 'Mbyte' else 1e9 if unit == 'Gbyte' else 1)
    elif 'bytes_write' in r[mn]:
        d['wr'] = v * (1e3 if unit == 'Kbyte' else 1e6 if unit == 'Mbyte' else 1e9 if unit == 'Gbyte' else 1)
    elif 'pipe_tensor' in r[mn]:
        d['tc'] = v
agg = collections.defaultdict(lambda: {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0, 'tc': 0.0})
for i, (k, d) in enumerate(launch.items()):
    if i < skip or 'us' not in d:
        continue
    name = d['name'].split('(')[0][-44:]
    a = agg[name]
    a['n'] += 1; a['us'] += d['us']; a['rd'] += d.get('rd', 0.0); a['wr'] += d.get('wr', 0.0); a['tc'] += d.get('tc', 0.0) * d['us']
tot = sum(a['us'] for a in agg.values())
print(f'sum of kernel times: {tot / 1e3:.2f} ms over {sum(a["n"] for a in agg.values())} launches (first {skip} skipped)')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    print(f'{name:44s} n={a["n"]:4d} {a["us"] / 1e3:8.3f} ms {100 * a["us"] / tot:5.1f}%  avg {a["us"] / a["n"]:8.1f} us  '
          f'dram rd {a["rd"] / a["n"] / 1e6:7.1f} wr {a["wr"] / a["n"] / 1e6:7.1f} MB/launch  tensor pipe {a["tc"] / max(a["us"], 1e-9):4.1f}%')

if len(sys.argv) > 3:
    import json
    js = {}
    for name, a in agg.items():
        short = name.split('::')[-1].split('<')[0].strip()
        js[short] = (a['rd'] + a['wr']) / a['n']
    json.dump(js, open(sys.argv[3], 'w'), indent=1, sort_keys=True)
