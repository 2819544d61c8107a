"""Print the per-launch-site table of a bench.py JSON line (forward.launch_sites)."""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], d["unit"], "ms/step", d["ms_per_step"], "forward ms", d["forward"]["ms"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_us", "share_of_step")})
tot = 0.0
for k, v in sorted(d["forward"]["launch_sites"].items(), key=lambda kv: -kv[1]["launches_per_forward"] * kv[1]["us_per_launch"]):
    ms = v["launches_per_forward"] * v["us_per_launch"] / 1e3
    tot += ms
    print(f"{ms:7.3f} ms  n={v['launches_per_forward']:5.1f}  {v['us_per_launch']:8.1f} us  {k}")
print(f"{tot:7.3f} ms total")
