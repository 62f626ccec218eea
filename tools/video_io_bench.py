"""Measurement tool: the script path's video output at 1080p.  (a) Stylization.transfer_stream alone (uint8 frames to the host),
(b) the same with every frame JPEG-encoded on the GPU while it is on the device (video_io.MjpgWriter through device_sink, bitstreams
muxed into an .avi), (c) what the reference does with the frames afterwards (generate_real_video.py:175-186): cv2.VideoWriter('MJPG')
on the host, per frame.  Prints one JSON line."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import torch

import bench
from rerevst_code_b200.framework import Stylization
from rerevst_code_b200.video_io import MjpgWriter
from rerevst_code_b200.weights import synthetic_state_dict

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
h, w = bench.SIZES["1080p"]
ph, pw = bench.padded_size(h, w)
fw = Stylization(synthetic_state_dict(0), cuda=True)
fw.prepare_style(bench.synthetic_frame(512, 512, 1))
fw.clean()
for i in range(2):
    fw.add(bench.synthetic_frame(h, w, 50 + i))
fw.compute()
raw = [bench.synthetic_frame(h, w, 100 + i) for i in range(4)]
tmp = tempfile.mkdtemp()


def run(sink=None):
    frames = (raw[i % 4] for i in range(N))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = None
    for out in fw.transfer_stream(frames, pad_to=(ph, pw), copy=False, out_dtype="u8", depth=4, device_sink=sink):
        last = out
    torch.cuda.synchronize()
    return N / (time.perf_counter() - t0), last.copy()


run()                                   # warm-up: graph capture
plain_fps, sample = run()
vw = MjpgWriter(os.path.join(tmp, "gpu.avi"), 24, (w, h), quality=75, states=4)


def sink(i, dev_frame, stream):
    k = i % 4
    vw.encode(dev_frame, k, stream)
    return lambda: vw.flush(k)


gpu_fps, _ = run(sink)
gpu_bytes = vw.bytes / max(1, vw.frames)
vw.release()
cw = cv2.VideoWriter(os.path.join(tmp, "cpu.avi"), cv2.VideoWriter_fourcc(*"MJPG"), 24, (w, h))
t0 = time.perf_counter()
for i in range(24):
    cw.write(sample)
cw.release()
cpu_fps = 24 / (time.perf_counter() - t0)
cap = cv2.VideoCapture(os.path.join(tmp, "gpu.avi"))
fr, ok = None, False
while True:                             # the last frame of the video is the frame `sample` holds
    got, f = cap.read()
    if not got:
        break
    fr, ok = f, True
print(json.dumps({"frames": N, "size": [h, w],
                  "transfer_stream_u8_fps": round(plain_fps, 1),
                  "transfer_stream_plus_gpu_mjpg_fps": round(gpu_fps, 1),
                  "gpu_mjpg_bytes_per_frame": int(gpu_bytes), "raw_u8_bytes_per_frame": h * w * 3,
                  "cv2_videowriter_mjpg_fps_one_host_thread": round(cpu_fps, 1),
                  "gpu_video_last_frame_psnr_vs_frame": round(cv2.PSNR(fr, sample) if ok else -1.0, 2),
                  "note": "stylized random-weight frames are noise-like (worst case for JPEG size and PSNR)"}))
