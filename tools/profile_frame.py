"""Dev/profiling tool: one stylized frame between cudaProfilerStart/Stop, for
    ncu --profile-from-start off [--set full | --metrics gpu__time_duration.sum] python tools/profile_frame.py
Everything before the marked frame (weight repack, style encoder, pre-pass, warm-up) is outside the capture."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from rerevst_code_b200.framework import Stylization
from rerevst_code_b200.weights import synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1080p")
ap.add_argument("--precision", default="x3")
ap.add_argument("--kernels", default="auto")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--mode", default="global", choices=["global", "frame"], help="frame: use_Global=False (style_network_frame.py)")
args = ap.parse_args()

h, w = bench.SIZES[args.size]
ph, pw = bench.padded_size(h, w)
fw = Stylization(synthetic_state_dict(0), cuda=True, precision=args.precision, impl=args.kernels, use_Global=args.mode == "global")
fw.prepare_style(bench.synthetic_frame(512, 512, 1))
if args.mode == "global":
    fw.clean()
    for i in range(2):
        fw.add(bench.synthetic_frame(h, w, 50 + i))
    fw.compute()
eng = fw.model._eng()
frames = [torch.from_numpy(bench.reflect_pad(bench.synthetic_frame(h, w, 100 + i), ph, pw)).unsqueeze(0).cuda() for i in range(2)]
post = ("f32", (64, 64, h, w))          # the finished frame, as bench.py times it
run = (lambda f: eng.forward(f, kind=1, post=post)) if args.mode == "global" else (lambda f: eng.forward_frame(f, kind=1))
for i in range(3):
    run(frames[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(args.frames):
    run(frames[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", args.frames, "frame(s) at", ph, "x", pw)
