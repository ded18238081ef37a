#!/usr/bin/env python
"""Per source line: instructions executed and stall samples of one kernel of an .ncu-rep
(needs --import-source on and -lineinfo).  usage: ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + rx, "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
acc = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":   # a source line summary row
        key = (cur_file, int(r[0]))
        i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples")
        src = r[1]
        a = acc.setdefault(key, [0, 0, src])
        a[0] += int(r[i_inst]) if r[i_inst].isdigit() else 0; a[1] += int(r[i_s]) if r[i_s].isdigit() else 0
tot_i = sum(a[0] for a in acc.values()) or 1
tot_s = sum(a[1] for a in acc.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
# by file
byf = collections.Counter(); bys = collections.Counter()
for (f, l), a in acc.items(): byf[f] += a[0]; bys[f] += a[1]
for f in byf: print(f"  {f}: inst {100*byf[f]/tot_i:.1f}%  samples {100*bys[f]/tot_s:.1f}%")
for (f, l), a in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:4d} inst {100*a[0]/tot_i:5.1f}% samp {100*a[1]/tot_s:5.1f}%  {a[2][:110]}")
