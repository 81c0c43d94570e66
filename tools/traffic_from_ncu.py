#!/usr/bin/env python3
"""profiles/r01_traffic.json from an `ncu --set full` report of the bench: mean
dram__bytes_read.sum + dram__bytes_write.sum per launch of each tick kernel.
usage: tools/traffic_from_ncu.py <report.ncu-rep> <legs> [out.json]"""
import csv
import io
import json
import subprocess
import sys

rep, legs = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/r01_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("wap::", "").replace("void ", "").split("<")[0].strip()
    b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
    acc.setdefault(name, []).append(b)
res = {}
for k, v in acc.items():
    m = sum(v) / len(v)
    res[k] = {"dram_bytes_per_launch_mean": m, "launches": len(v), "legs": legs, "dram_bytes_per_leg_frame": m / legs}
res["_source"] = "ncu --set full (%s), %d legs, steady state (ticks >= 300; 2- and 3-block); tools/traffic_from_ncu.py" % (rep, legs)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
