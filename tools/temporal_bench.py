"""Dev/measurement tool for BASELINE config 5 (train.py:375-388 temporal-loss path): B=4 pairs at 512x512, synthetic flow of
GenerateFakeFlow's range (+-16 px).  Times warp, TemporalLoss.forward (fused warp + L1 mean), the Vgg19 loss features of two
styled frames and validation(); prints achieved HBM GB/s of the warp against MEASURED_PEAKS.json.
usage: python tools/temporal_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from rerevst_code_b200.loss_networks import TemporalLoss, warp
from rerevst_code_b200.style_networks import TransformerNet
from rerevst_code_b200.weights import synthetic_state_dict

dev = torch.device("cuda", 0)
B, C, H, W = 4, 3, 512, 512
g = torch.Generator().manual_seed(0)
first = torch.randn(B, C, H, W, generator=g).to(dev)
flow = (torch.randn(B, 2, H // 64, W // 64, generator=g) * 8).to(dev)
flow = torch.nn.functional.interpolate(flow, size=(H, W), mode="bilinear", align_corners=False).contiguous()
second = warp(first, flow) + 1e-3 * torch.randn(B, C, H, W, generator=g).to(dev)
tl = TemporalLoss()
net = TransformerNet().to(dev)
net.load_state_dict(synthetic_state_dict(0))
style = torch.randn(1, 3, 256, 256, generator=g).to(dev)


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
with torch.no_grad():
    ms_warp = timed(lambda: warp(first, flow))
    ms_tl = timed(lambda: tl(first, second, flow))
    ms_vgg = timed(lambda: (net.vgg19(first), net.vgg19(second)), iters=10, warm=2)
    ms_val = timed(lambda: net.validation(first[:1], style), iters=5, warm=1)
bytes_warp = B * H * W * (8 + 4 * C + 4 * C)          # flow + gather + write (32 B per pixel for C = 3)
bytes_tl = B * H * W * (8 + 4 * C + 4 * C + 4 * C)    # + the second frame
print(json.dumps({
    "config": "B=4, 512x512, C=3, flow +-16 px (BASELINE configs[4])",
    "warp_ms": ms_warp, "warp_gbs": bytes_warp / ms_warp / 1e6, "warp_frac_of_hbm_peak": bytes_warp / ms_warp / 1e6 / peaks["hbm_gbs"],
    "temporal_loss_ms": ms_tl, "temporal_loss_gbs": bytes_tl / ms_tl / 1e6,
    "vgg19_two_batches_ms": ms_vgg, "vgg19_tflops_algorithmic": 2 * B * 126.53e9 / (ms_vgg * 1e-3) / 1e12,
    "validation_one_frame_ms": ms_val,
    "note": "8.4 MB per warp call: launch-latency bound, not HBM bound (SURVEY 8d)"}))
