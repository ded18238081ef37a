#!/usr/bin/env python
"""SASS evidence of the sm_100a build: per kernel, counts of the instructions the design relies on
(cuobjdump -sass; no GPU needed).  usage: python tools/sass_grep.py > profiles/rNN_sass_grep.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sings_b200", "lib", "libsings_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("REDGx4", r"\bRED\S*\.F32x4|REDG\S*F32x4|RED\.E\.ADD\.F32x4"),
        ("FFMA2", r"\bFFMA2"), ("FMUL2", r"\bFMUL2"), ("FADD2", r"\bFADD2"), ("PREEXIT", r"\bPREEXIT"),
        ("MMA", r"\b(HMMA|IMMA|DMMA|QMMA|UTCHMMA|UTCQMMA|UTCIMMA|UTCMMA)")]
print("# SASS evidence of the sm_100a build (cuobjdump -sass sings_b200/lib/libsings_b200.so)")
print("# per kernel: instruction counts of the Blackwell / Hopper+ features the design relies on")
print("#   UBLKCP = cp.async.bulk (TMA 1-D bulk copy)   SYNCS = mbarrier operations   LDGSTS = cp.async (global -> shared)")
print("#   REDGx4 = red.global.add.v4.f32 (vector reduction)   FFMA2 / FMUL2 / FADD2 = packed binary32 pairs (fma/mul/add.rn.f32x2)")
print("#   PREEXIT = griddepcontrol.launch_dependents (programmatic dependent launch)   MMA = any tensor-core instruction (none: see DESIGN.md 5)")
print(f"{'kernel':62s}{'instr':>6s}" + "".join(f"{c:>8s}" for c, _ in cols))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, name in zip(blocks, names):
    body = [l for l in blk.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    short = re.sub(r"\((int|bool)\)", "", name).replace("void ", "")
    short = re.sub(r"\(.*", "", short)
    print(f"{short[:61]:62s}{len(body):6d}" + "".join(f"{sum(1 for l in body if re.search(rx, l)):8d}" for _, rx in cols))
