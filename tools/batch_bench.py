"""Dev tool: frames/s of the graphed per-frame forward when B frames go through one launch sequence
(wave quantisation of the low-resolution layers).  usage: python tools/batch_bench.py [B ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rerevst_code_b200.framework import Stylization
from rerevst_code_b200.weights import synthetic_state_dict

h, w = bench.SIZES["1080p"]
ph, pw = bench.padded_size(h, w)
fw = Stylization(synthetic_state_dict(0), cuda=True, precision=os.environ.get("RRV_PRECISION", "x3"))
fw.prepare_style(bench.synthetic_frame(512, 512, 1))
fw.clean()
for i in range(2):
    fw.add(bench.synthetic_frame(h, w, 50 + i))
fw.compute()
eng = fw.model._eng()
frames = [torch.from_numpy(bench.reflect_pad(bench.synthetic_frame(h, w, 100 + i), ph, pw)).unsqueeze(0).cuda() for i in range(4)]
for B in [int(a) for a in sys.argv[1:]] or [1, 2]:
    batches = [torch.cat([frames[(i + j) % 4] for j in range(B)], 0) for i in range(4)]
    for i in range(3):
        eng.forward_graphed(batches[i % 4], kind=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for i in range(steps):
        eng.forward_graphed(batches[i % 4], kind=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"B={B}: {ms:.3f} ms per launch sequence, {ms / B:.3f} ms per frame, {1e3 * B / ms:.1f} frames/s", flush=True)
