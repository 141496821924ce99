"""Aggregate the per-instruction stall samples of an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source sass`)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[2].isdigit()]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S = sum(int(r[ix['# Samples']]) for r in body)
tot = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print('samples', S, 'instructions', len(body))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f'  {s:26s} {v:9d} {v / S:.3f}')
for i, r in enumerate(body):
    r.append(i)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(body, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = sorted(((s, int(r[ix[s]])) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{r[-1]:5d} {int(r[ix['# Samples']]):7d} {int(r[ix['Instructions Executed']]):10d}  {r[ix['Source']].strip()[:60]:60s} {st}")
