"""BASELINE.json configs[3] as written: a 1024-frame 1080p clip, frame-parallel over the ranks of one node, the 128 sampled
frames of its pre-pass (generate_real_video.py:133-143: every 8th frame + the last one) sharded over the ranks with NCCL
all-gathers of the mergeable per-channel partials (rerevst-code_b200/dist.py).  STRONG scaling: the clip is fixed, each rank
stylizes frames [r * 1024 / G, (r + 1) * 1024 / G).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/config4.py
    python tools/config4.py --frames 256            # one GPU: BASELINE configs[2]'s clip (32 samples) with the fp32-accurate kernels

Rank 0 prints one JSON line: pre-pass seconds (max over ranks) and peak device memory per rank, frames/s of the clip loop
device-resident and end to end (host uint8 frames in, host uint8 frames out), and that every rank holds bit-identical tables.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench


def quick_frame(h, w, seed):
    """Image-like uint8 frame, cheap enough to make 128 distinct sampled frames on the host (bench.synthetic_frame is ~1 s each)."""
    rng = np.random.RandomState(seed)
    coarse = (rng.rand(h // 32 + 2, w // 32 + 2, 3) * 255).astype(np.float32)
    img = np.kron(coarse, np.ones((32, 32, 1), np.float32))[:h, :w]
    img = (img + np.roll(img, 11, 0) + np.roll(img, 13, 1) + np.roll(img, (5, 7), (0, 1))) * 0.25
    return np.clip(img + rng.randint(-6, 7, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--distinct", type=int, default=8, help="distinct host frames per rank rotated through the clip loop")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rerevst_code_b200.dist import sample_indices, shard_range, sharded_prepass
    from rerevst_code_b200.framework import Stylization
    from rerevst_code_b200.weights import synthetic_state_dict

    h, w = bench.SIZES[args.size]
    ph, pw = bench.padded_size(h, w)
    fw = Stylization(synthetic_state_dict(0), cuda=True)
    fw.prepare_style(bench.synthetic_frame(512, 512, 1))
    idx = sample_indices(args.frames)                       # frames of the clip the script samples
    lo, hi = shard_range(len(idx), rank, world)

    # every rank indexes the same list; only its shard (+ the clip's first sample, quirk Q1) is synthesised, BEFORE the timed
    # region: making the frames on the host is not pre-pass time
    mine = {i: quick_frame(h, w, 1000 + idx[i]) for i in set(range(lo, hi)) | {0}}

    class Lazy(list):
        def __getitem__(self, i):
            if isinstance(i, slice):
                return [self[j] for j in range(*i.indices(len(self)))]
            return mine[i]
    samples = Lazy([None] * len(idx))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    sharded_prepass(fw, samples, rank, world)
    torch.cuda.synchronize()
    pre_s = time.perf_counter() - t0
    peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    eng = fw.model._eng()
    st = eng.export_clip_state()
    digest = torch.cat([v.flatten() for v in st["stats"].values()] + [t.flatten() for ab in st["filters"].values() for t in ab])

    f_lo, f_hi = shard_range(args.frames, rank, world)
    n_local = f_hi - f_lo
    host = [bench.reflect_pad(quick_frame(h, w, 5000 + rank * 64 + i), ph, pw) for i in range(args.distinct)]
    devf = [torch.from_numpy(f).unsqueeze(0).to(dev) for f in host]
    crop = (64, 64, h, w)
    post = ("u8", crop)
    for i in range(4):
        eng.forward_graphed(devf[i % len(devf)], kind=1, post=post)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_local):
        eng.forward_graphed(devf[i % len(devf)], kind=1, post=post)
    e1.record()
    sync()
    ms_dev = gmax(e0.elapsed_time(e1))
    list(fw.transfer_stream((host[i % len(host)] for i in range(6)), crop=crop, copy=False, out_dtype="u8"))
    sync()
    t0 = time.perf_counter()
    n = 0
    for res in fw.transfer_stream((host[i % len(host)] for i in range(n_local)), crop=crop, copy=False, out_dtype="u8"):
        n += int(res[0, 0, 0]) >= 0
    torch.cuda.synchronize()
    ms_e2e = gmax((time.perf_counter() - t0) * 1e3)
    pre_s, peak_gb = gmax(pre_s), gmax(peak_gb)
    same = True
    if world > 1:
        all_d = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(all_d, digest)
        same = all(torch.equal(all_d[0], d) for d in all_d)
    if rank == 0:
        print(file=bench._JSON_OUT, flush=True, *[json.dumps({
            "config": f"{args.frames}-frame {args.size} clip padded to {ph}x{pw}, {len(idx)} sampled frames ({h}x{w}, unpadded) sharded over "
                      f"{world} rank(s): {hi - lo} per rank, strong scaling ({n_local} frames per rank)",
            "n_gpus": world, "prepass_s_max_over_ranks": pre_s, "prepass_peak_device_memory_gb_max_over_ranks": peak_gb,
            "tables_bit_identical_on_all_ranks": bool(same),
            "clip_loop_device_frames_per_s": args.frames / (ms_dev * 1e-3), "clip_loop_device_s": ms_dev * 1e-3,
            "clip_loop_e2e_frames_per_s": args.frames / (ms_e2e * 1e-3), "clip_loop_e2e_s": ms_e2e * 1e-3,
            "e2e": "host uint8 frames in, host uint8 BGR frames out (transfer_stream, out_dtype='u8')",
            "whole_job_s": pre_s + ms_e2e * 1e-3})])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
