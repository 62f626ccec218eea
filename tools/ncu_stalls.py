"""Dev tool: stall-reason totals and the hottest SASS instructions of one kernel of a .ncu-rep (source page).
usage: python tools/ncu_stalls.py REP LAUNCH_INDEX [TOP]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:110])
h = rows[1]
si, src = h.index("# Samples"), h.index("Source")
body = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = {}
for r in body:
    for c in stall_cols:
        if r[c].isdigit():
            agg[h[c]] = agg.get(h[c], 0) + int(r[c])
print("samples", tot, " ".join(f"{k[6:]}={100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
order = {id(r): i for i, r in enumerate(body)}
for r in sorted(sorted(body, key=lambda r: -int(r[si]))[:top_n], key=lambda r: order[id(r)]):
    st = sorted(((int(r[c]), h[c][6:]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:2]
    print(f"{order[id(r)]:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:4.1f}% {r[src].strip()[:86]:86s} {st}")
