"""Dev tool: one-line summary of a bench.py JSON line read from stdin."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d.get("roofline", {})
print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 4), "frac", round(r.get("frac", 0), 4),
      "kernel_ms", round(r.get("kernel_ms_per_frame", 0), 4), "parity", (d.get("parity") or {}).get("ok"), d.get("clocks"))
if "--layers" in sys.argv:
    for l in d.get("layers", []):
        print(f"  {l['layer']:36s} {l['ms']:.4f} {l['tflops']:.0f}")
