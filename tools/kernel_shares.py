"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r"\(.*", "", r[ik]).replace("m0::", "")
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / tot * 100:5.1f}%  n={n:4d}  avg={t / n / 1e3:10.1f} us  {k}")
