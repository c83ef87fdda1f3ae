"""Per kernel-class time of one device-resident page (uses tools/layer_times.py's profile JSON on its last line)."""
import json
import subprocess
import sys
out = subprocess.run([sys.executable, "tools/layer_times.py"], capture_output=True, text=True).stdout.strip().splitlines()
d = json.loads(out[-1])
tot = 0.0
for k, v in sorted(d.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:18s} {v['launches']:5d} {v['ms']:8.3f} ms")
    tot += v["ms"]
print(f"{'total':18s}       {tot:8.3f} ms")
