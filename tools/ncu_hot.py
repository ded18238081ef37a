"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K`:
top SASS instructions by stall samples, with the dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the file may hold several launches: keep the first block only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"instructions in kernel: {len(body)}; executed warp-instr: {inst}; samples: {tot}")
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[ix["# Samples"]] or 0)
    why = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f"{i:5d} {n:6d} {100*n/max(tot,1):5.1f}%  x{r[ix['Instructions Executed']]:>8s}  {why[6:]:14s} {r[ix['Source']].strip()[:90]}")
