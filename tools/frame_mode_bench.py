"""Dev tool: frames/s of the per-frame mode (use_Global=False, style_network_frame.py) at 1080p padded, device-resident uint8 frames.
usage: python tools/frame_mode_bench.py [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rerevst_code_b200.framework import Stylization
from rerevst_code_b200.weights import synthetic_state_dict

size = sys.argv[1] if len(sys.argv) > 1 else "1080p"
h, w = bench.SIZES[size]
ph, pw = bench.padded_size(h, w)
fw = Stylization(synthetic_state_dict(0), cuda=True, use_Global=False)
fw.prepare_style(bench.synthetic_frame(512, 512, 1))
eng = fw.model._eng()
frames = [torch.from_numpy(bench.reflect_pad(bench.synthetic_frame(h, w, 100 + i), ph, pw)).unsqueeze(0).cuda() for i in range(2)]
for i in range(3):
    eng.forward_frame_graphed(frames[i % 2], kind=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 40
e0.record()
for i in range(steps):
    eng.forward_frame_graphed(frames[i % 2], kind=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"frame mode {size} ({ph}x{pw}): {ms:.2f} ms per frame, {1e3 / ms:.1f} frames/s")
