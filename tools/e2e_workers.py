"""e2e pages/s of the dev bench for several host worker-thread counts (development tool)."""
import json
import subprocess
import sys
for w in sys.argv[1:]:
    out = subprocess.run([sys.executable, "bench.py", "--pages", "16", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--fast-e2e",
                          "--workers", w], capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    print("workers", w, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), flush=True)
