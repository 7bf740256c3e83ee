"""Hot spots of one kernel of an .ncu-rep source page (SASS view): the instructions with the most stall samples, and the
sample distribution over windows of the instruction stream.
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv;  python tools/ncu_hot.py src.csv <section> [top] [window]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
k = int(sys.argv[2]); ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30; win = int(sys.argv[4]) if len(sys.argv) > 4 else 400
print(rows[secs[k]][1][:100])
H = rows[secs[k] + 1]
isrc, isamp, iex = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
sc = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
data = [(r[isrc].strip(), int(r[isamp]), int(r[iex]), [int(r[i]) for i in sc]) for r in rows[secs[k] + 2:secs[k + 1]] if len(r) >= len(H)]
tot = sum(d[1] for d in data)
print('instructions', len(data), 'samples', tot, 'warp-instr executed', sum(d[2] for d in data))
for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][1])[:ntop]):
    s, n, ex, st = data[i]
    dom = max(range(len(st)), key=lambda j: st[j])
    print(f'{i:6d} {100 * n / tot:5.2f}% ex={ex:9d} {H[sc[dom]][6:]:18s} {s[:80]}')
for w in range(0, len(data), win):
    n = sum(d[1] for d in data[w:w + win]); ex = sum(d[2] for d in data[w:w + win])
    st = [sum(d[3][j] for d in data[w:w + win]) for j in range(len(sc))]
    o = sorted(range(len(st)), key=lambda j: -st[j])[:3]
    print(f'[{w:6d},{w + win:6d}) {100 * n / tot:5.1f}% exec {ex:10d}  ' + ', '.join(f'{H[sc[j]][6:]}={st[j]}' for j in o))
