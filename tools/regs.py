"""Registers / spills / static shared memory per kernel from the ptxas logs of the last build."""
import glob, re, subprocess
for f in sorted(glob.glob('sings_b200/_build/*.ptxas.log')):
    name = None
    for l in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            name = subprocess.run(['c++filt', '-p', m.group(1)], capture_output=True, text=True).stdout.strip()
        m = re.search(r'(\d+) bytes spill stores', l)
        if m: spill = int(m.group(1))
        m = re.search(r'Used (\d+) registers.*?(?:, (\d+) bytes smem)?', l)
        if m and name:
            sm = re.search(r'(\d+) bytes smem', l)
            print(f"{name[:70]:70s} regs={m.group(1):>3s} spill={spill:3d} smem={sm.group(1) if sm else 0}")
            name = None
