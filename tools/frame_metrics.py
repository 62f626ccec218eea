"""Dev tool: per-launch table (time, DRAM bytes, L2->SM bytes, tensor-pipe %) from an ncu --csv metrics log of one frame;
writes profiles/r1_traffic.json (bench.py's roofline.traffic comes from it)."""
import csv
import json
import sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, mi, vi, ui, ii = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "%": 1.0}
L = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    d = L.setdefault(int(r[ii]), {"k": r[ki].split("(")[0][-28:]})
    try:
        d[r[mi]] = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    except ValueError:
        d[r[mi]] = float("nan")
T, DR, DW, XB, TP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
                     "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
print(f"{'#':>3s} {'kernel':30s} {'us':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'L2->SM MB':>10s} {'tensor %':>8s}")
tot = dict(t=0.0, d=0.0)
conv = dict(t=0.0, d=0.0, n=0)
for i in sorted(L):
    d = L[i]
    print(f"{i:3d} {d['k']:30s} {d[T]:8.1f} {d[DR]:9.1f} {d[DW]:9.1f} {d[XB]:10.1f} {d[TP]:8.1f}")
    tot["t"] += d[T]
    tot["d"] += d[DR] + d[DW]
    if "conv_tc" in d["k"]:
        conv["t"] += d[T]
        conv["d"] += d[DR] + d[DW]
        conv["n"] += 1
print(f"frame: {tot['t']:.1f} us, {tot['d']:.1f} MB DRAM; conv kernels: {conv['n']} launches, {conv['t']:.1f} us, {conv['d']:.1f} MB DRAM")
if len(sys.argv) > 2:
    json.dump({"source": f"{src} (ncu --clock-control none, one 1216x2048 frame, x3 mode; per-launch values are cold-cache and serialised)",
               "conv_launches_per_frame": conv["n"], "conv_dram_bytes_per_frame": conv["d"] * 1e6,
               "frame_dram_bytes": tot["d"] * 1e6}, open(sys.argv[2], "w"), indent=1)
