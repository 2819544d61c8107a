"""Write profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernels) from an `ncu --set full` report:
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep
bench.py reads the file for roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, units, data = rows[0], rows[1], rows[2:]
ik, ir, iw, it = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in data:
    name = r[ik].split("(")[0].replace("void ", "").replace("m0::tc::", "").replace("m0::", "").split("<")[0]
    b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += b
    e["time_us"] += float(r[it].replace(",", "")) * (1e-3 if units[it] == "ns" else 1.0 if units[it] in ("us", "usecond") else 1e3)
for e in out.values():
    e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
    e["time_us_per_launch"] = e["time_us"] / e["launches"]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
for e in out.values():
    e["source"] = os.path.basename(rep)
if os.path.exists(path):                       # keep the entries of kernels captured in other reports
    old = {k: v for k, v in json.load(open(path)).items() if isinstance(v, dict)}
    old.update(out)
    out = old
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
