"""stdin: bench.py output; prints one line: label, device-timed ms/step, e2e ms/step."""
import json, sys
label = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(label, "dev", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "value", round(d["value"] / 1e6, 1), d["unit"])
