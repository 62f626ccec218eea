"""Dev tool: top stall-sampled SASS instructions of a .ncu-rep (source page), with their dominant stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si = h.index("# Samples")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
body = rows[2:]
tot = sum(int(r[si]) for r in body if len(r) > si and r[si].isdigit())
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si]) if body[i][si].isdigit() else 0)[:top]
print("total samples", tot)
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c]), h[c]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[si]):7d} {100*int(r[si])/tot:5.1f}%  {r[1].strip()[:70]:70s} {st}")
