"""profiles/<tag>_ncu_compact.txt (tools/ncu_compact.py) -> profiles/ncu_traffic.json: DRAM bytes per
launch (dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full` capture) per kernel.
bench.py reads the JSON to fill `roofline.traffic`.   usage: python tools/ncu_traffic.py <compact.txt>"""
import json, os, re, sys
src = sys.argv[1]
out = {"source": os.path.basename(src), "kernels": {}}
for line in open(src):
    m = re.match(r"(.*?) us=([\d.e+-]+) rdMB=([\d.e+-]+) wrMB=([\d.e+-]+)", line)
    if not m:
        continue
    name = m.group(1).replace("void ", "").split("<")[0].strip()
    k = out["kernels"].setdefault(name, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    k["launches"] += 1
    k["us"] += float(m.group(2))
    k["dram_bytes"] += (float(m.group(3)) + float(m.group(4))) * 1e6
for k in out["kernels"].values():
    n = k.pop("launches")
    k["us"] = round(k["us"] / n, 2)
    k["dram_bytes"] = int(k["dram_bytes"] / n)
json.dump(out, open(os.path.join(os.path.dirname(src), "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
