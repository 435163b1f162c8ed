"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
keeps the LAST forward pass (from the last patch_embed_ln / resnet_stem launch on) and prints per-kernel totals (time share,
DRAM bytes when captured).  `--order` also lists the launches in order; `--traffic-json PATH` writes the mean DRAM
bytes per gemm_tc launch (bench.py's roofline.traffic)."""
import collections, csv, json, sys
args = [a for a in sys.argv[1:] if not a.startswith('--')]
path = args[0]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; ID, kn, mn, mv, mu = (h.index(x) for x in ('ID', 'Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
grid = h.index('Grid Size')
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    d = launch.setdefault(r[ID], {'name': r[kn].split('(')[0].replace('void ', ''), 'grid': r[grid]})
    v = float(r[mv].replace(',', ''))
    if r[mn] == 'gpu__time_duration.sum': d['us'] = v / 1e3 if r[mu] == 'ns' else v * 1e3 if r[mu] == 'ms' else v
    else:
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[mu], 1)
        d[r[mn]] = v * scale
L = list(launch.values())
start = max(i for i, d in enumerate(L) if 'patch_embed_ln' in d['name'] or 'resnet_stem' in d['name'])
L = L[start:]
tot = sum(d['us'] for d in L)
agg = collections.OrderedDict()
for d in L:
    a = agg.setdefault(d['name'], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d['us']
    a[2] += d.get('dram__bytes_read.sum', 0.0); a[3] += d.get('dram__bytes_write.sum', 0.0)
has_dram = any('dram__bytes_read.sum' in d for d in L)
print(f'last forward pass: {len(L)} launches, {tot:.1f} us (cold-cache, serialised)')
for k, (c, t, br, bw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f'  dram rd {br/1e6:8.1f} MB  wr {bw/1e6:8.1f} MB  ({(br+bw)/t/1e6:6.2f} TB/s)' if has_dram else ''
    print(f'{t:10.1f} us {100*t/tot:5.1f}%  x{c:4d}  {k}{extra}')
if has_dram:
    print(f'step DRAM traffic: read {sum(a[2] for a in agg.values())/1e9:.2f} GB, write {sum(a[3] for a in agg.values())/1e9:.2f} GB')
for a in sys.argv[1:]:
    if a.startswith('--traffic-json'):
        out = a.split('=', 1)[1]
        g = [d for d in L if 'gemm_tc_kernel' in d['name']]
        json.dump({'source': path.split('/')[-1], 'kernel': 'gemm_tc_kernel (all variants)', 'launches': len(g),
                   'dram_bytes_per_launch': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in g) / len(g),
                   'dram_bytes_per_step': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in g),
                   'note': 'ncu serialised launches, cold L2 per launch replay: an upper bound of the in-graph traffic'}, open(out, 'w'), indent=1)
if '--order' in sys.argv:
    print('--- launches in order')
    for d in L: print(f"{d['us']:9.1f} us  {d['grid']:>14s}  {d['name']}")
