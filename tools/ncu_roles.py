"""Dev tool: split the stall samples of a conv_tc2 kernel capture by warp role (TMA producer / MMA issuer / epilogue),
using the distinctive SASS of each role as region markers.  usage: python tools/ncu_roles.py rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", sys.argv[2]] if len(sys.argv) > 2 else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si = h.index("# Samples")
body = [r for r in rows[2:] if len(r) > si]
marks = []
for i, r in enumerate(body):
    s = r[1]
    if "UTMALDG" in s:
        marks.append((i, "tma"))
    elif "UTCHMMA" in s or "UTCMMA" in s or "UTCBAR" in s:
        marks.append((i, "mma"))
    elif "LDTM" in s:
        marks.append((i, "epi"))
print("markers:", [(i, k) for i, k in marks][:3], "...", len(marks))
# region boundaries: midpoints between the last marker of one kind and the first of the next
first = {}
last = {}
for i, k in marks:
    first.setdefault(k, i)
    last[k] = i
print("first/last:", first, last)
tot = sum(int(r[si]) for r in body if r[si].isdigit())
def region(lo, hi):
    return sum(int(body[i][si]) for i in range(lo, hi) if body[i][si].isdigit())
order = sorted(first, key=lambda k: first[k])
bounds = [0]
for a, b in zip(order, order[1:]):
    bounds.append((last[a] + first[b]) // 2)
bounds.append(len(body))
print("total samples", tot)
for k, lo, hi in zip(order, bounds, bounds[1:]):
    n = region(lo, hi)
    print(f"  {k}: instr {lo}..{hi}  samples {n} ({100 * n / tot:.1f}%)")
    top = sorted(range(lo, hi), key=lambda i: -int(body[i][si]) if body[i][si].isdigit() else 0)[:6]
    for i in sorted(top):
        print(f"      {i:5d} {body[i][si]:>6s}  {body[i][1].strip()[:80]}")
