"""Registers / spills / smem per kernel from `nvcc -Xptxas -v` output (stdin or a log file)."""
import re, subprocess, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
name = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: spill = m.group(1)
    m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        print(f"{name[:60]:60s} regs={m.group(1):>3s} spill={spill} smem={sm.group(1) if sm else 0}")
        name = None
