"""Quick e2e check: C2 bench line's value / e2e and kernel times (short)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-secondary", "--no-cpu-baseline"], capture_output=True, text=True)
for l in out.stdout.splitlines():
    if l.startswith('{"metric'):
        d = json.loads(l)
        print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "e2e p50", round(d["e2e"]["latency_ms"]["p50"], 3), d["kernels_ms"])
print(out.stderr[-2000:])
