"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ci = {k: i for i, k in enumerate(h)}
S = ci['Warp Stall Sampling (All Samples)']; src = ci['Source']; ex = ci['Instructions Executed']
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
body = rows[2:]
tot = sum(int(r[S]) for r in body if r[S].isdigit())
print('total samples', tot)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
idx = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -(int(r[S]) if r[S].isdigit() else 0))[:n]:
    reasons = sorted(((int(r[ci[k]]), k[6:]) for k in stall_cols if r[ci[k]].isdigit() and int(r[ci[k]]) > 0), reverse=True)[:3]
    print(f"{100*int(r[S])/tot:5.1f}%  line {idx[id(r)]:4d} exec {r[ex]:>8s}  {r[src].strip()[:70]:70s} {reasons}")
