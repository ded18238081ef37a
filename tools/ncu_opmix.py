"""Executed warp-instructions by opcode for one kernel: ncu -i X.ncu-rep --page source --csv --kernel-name K | python tools/ncu_opmix.py"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "RED", "ATOM")) else op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n; tot += n
print("total warp-instr", tot)
for op, n in ops.most_common(28): print(f"{op:14s} {n:10d} {100*n/tot:5.1f}%")
