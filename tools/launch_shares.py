"""Per-kernel average device time and share from an ncu launch-list CSV."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
h = rows[0]; i_n = h.index('Kernel Name'); i_v = h.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) > i_v:
        try: d[r[i_n].split('(')[0]].append(float(r[i_v].replace(',', '')))
        except ValueError: pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:50]:50s} n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f}us share={sum(v)/tot*100:5.1f}%")
