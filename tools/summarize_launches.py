"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: keeps the LAST forward pass (from the last
patch_embed_ln_kernel launch on) and prints per-kernel totals + the GEMM launches in order."""
import csv, collections, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
grid = h.index('Grid Size') if 'Grid Size' in h else None
L = []
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', '')); v = v / 1e3 if r[mu] == 'ns' else v * 1e3 if r[mu] == 'ms' else v
    L.append((r[kn].split('(')[0].replace('void ', ''), v, r[grid] if grid is not None else ''))
start = max(i for i, (n, _, _) in enumerate(L) if 'patch_embed_ln' in n)
L = L[start:]
agg = collections.OrderedDict(); tot = sum(v for _, v, _ in L)
for n, v, _ in L:
    d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += v
print(f'last forward pass: {len(L)} launches, {tot:.1f} us (cold-cache, serialised)')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f'{t:10.1f} us {100*t/tot:5.1f}%  x{c:4d}  {k}')
if len(sys.argv) > 2:
    print('--- launches in order'); [print(f'{v:9.1f} us  {g:>12s}  {n}') for n, v, g in L]
