"""Dev tool: algorithmic DRAM bytes of every launch of one 1216x2048 frame (x3 mode: every activation is two 16-bit planes,
4 bytes per element; each layer reads its input once, its residual once, writes its output once) next to the measured
dram__bytes of profiles/r1_frame_metrics.csv.  usage: python tools/traffic_model.py [csv]"""
import csv
import sys

H, W = 1216, 2048
src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r1_frame_metrics.csv"
MB = 1e6
px = lambda s: (H // s) * (W // s)
E = 4                                                    # bytes per activation element in x3 mode (hi + lo)
layers = [("first layer (u8 in)", px(1) * 3, px(1) * 64 * E)]
layers += [("conv1_2 + pool", px(1) * 64 * E, px(2) * 64 * E), ("conv2_1", px(2) * 64 * E, px(2) * 128 * E),
           ("conv2_2 + pool", px(2) * 128 * E, px(4) * 128 * E), ("conv3_1", px(4) * 128 * E, px(4) * 256 * E),
           ("conv3_2", px(4) * 256 * E, px(4) * 256 * E), ("conv3_3", px(4) * 256 * E, px(4) * 256 * E),
           ("conv3_4 + pool", px(4) * 256 * E, px(8) * 256 * E), ("conv4_1 + norm0", px(8) * 256 * E, px(8) * 512 * E)]
for i in range(3):
    layers += [(f"Filter{i + 1}.down", px(8) * 512 * E, px(8) * 32 * E), (f"Filter{i + 1}.up + res", px(8) * (32 + 512) * E, px(8) * 512 * E)]
for name, s, cin, cout in (("slice4", 8, 512, 256), ("slice3", 4, 256, 128), ("slice2", 2, 128, 64)):
    layers += [(f"{name}.shortcut (fp32 out)", px(s) * cin * E, px(s) * cout * 4),
               (f"{name}.conv1 (x2)", px(s) * cin * E, px(s // 2) * cout * E),
               (f"{name}.conv2 + res", px(s // 2) * cout * E + px(s) * cout * 4, px(s // 2) * cout * E)]
layers += [("slice1 (RGB head, fp32 out)", px(1) * 64 * E, px(1) * 3 * 4)]

rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, mi, vi, ui, ii = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
meas = {}
for r in rows[hdr + 1:]:
    if len(r) > vi and r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        meas.setdefault(int(r[ii]), {})[r[mi]] = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
ids = sorted(meas)
print(f"{'launch':32s} {'model R':>9s} {'meas R':>9s} {'model W':>9s} {'meas W':>9s}   (MB)")
tm = tr = 0.0
for (name, rb, wb), i in zip(layers, ids):
    m = meas[i]
    print(f"{name:32s} {rb / MB:9.1f} {m['dram__bytes_read.sum']:9.1f} {wb / MB:9.1f} {m['dram__bytes_write.sum']:9.1f}")
    tm += (rb + wb) / MB
    tr += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
print(f"frame: model {tm:.0f} MB, measured {tr:.0f} MB ({tr / tm:.2f}x)")
