"""Host logic of the frame-parallel driver (rerevst-code_b200/dist.py): sharding, the all-gather of
mergeable partial statistics over gloo (world_size 2, CPU), and -- on GPUs -- the sharded pre-pass."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything_once():
    from rerevst_code_b200.dist import shard_range
    for n in (1, 2, 7, 8, 33, 128, 1024):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_sample_indices_match_reference_script():
    """generate_real_video.py:133-143: every 8th frame for s in range((n-1)//8), then the last frame."""
    from rerevst_code_b200.dist import sample_indices
    assert sample_indices(64) == [0, 8, 16, 24, 32, 40, 48, 63]
    assert sample_indices(33) == [0, 8, 16, 24, 32]
    assert len(sample_indices(256)) == 32 and len(sample_indices(1024)) == 128
    assert sample_indices(1) == [0]


def _partials(x):
    """numpy restatement of rrv_channel_stats: {count, sum, M2 about the local mean, min, max} per channel."""
    x = x.reshape(-1, x.shape[-1]).astype(np.float64)
    n = np.full(x.shape[1], x.shape[0], np.float64)
    s = x.sum(0)
    m2 = ((x - s / n) ** 2).sum(0)
    return np.stack([n, s, m2, x.min(0), x.max(0)])


def _chan_merge(parts):
    """Chan/Golub/LeVeque left-to-right merge (csrc/stats.cu: stats_merge_kernel)."""
    n, s, m2, mn, mx = parts[0]
    for p in parts[1:]:
        delta = p[1] / p[0] - s / n
        m2 = m2 + p[2] + delta * delta * n * p[0] / (n + p[0])
        n, s = n + p[0], s + p[1]
        mn, mx = np.minimum(mn, p[3]), np.maximum(mx, p[4])
    return np.stack([n, s, m2, mn, mx])


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rerevst_code_b200.dist import allgather_parts, shard_range
    rng = np.random.RandomState(0)
    samples = rng.randn(5, 6, 7, 16) * 2 + 0.5                 # 5 sampled frames, NHWC, every rank sees the same
    lo, hi = shard_range(5, rank, world)
    part = torch.from_numpy(_partials(samples[lo:hi]))
    parts = allgather_parts(part)                               # [world, 5, C] in rank order
    merged = _chan_merge([p.numpy() for p in parts])
    q.put((rank, tuple(parts.shape), merged))
    dist.destroy_process_group()


def test_allgather_and_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    whole = _partials(rng.randn(5, 6, 7, 16) * 2 + 0.5)
    for rank, shape, merged in got:
        assert shape == (2, 5, 16)
        assert np.array_equal(merged[0], whole[0]) and np.array_equal(merged[3:], whole[3:])
        assert np.allclose(merged[1], whole[1], rtol=1e-12) and np.allclose(merged[2], whole[2], rtol=1e-10)
    assert np.array_equal(got[0][2], got[1][2])                # bit-identical on every rank (fixed merge order)


# ------------------------------------------------------------------------------------------ GPU

def _nccl_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank % torch.cuda.device_count())
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    from rerevst_code_b200.dist import sharded_prepass
    from rerevst_code_b200.framework import Stylization
    from rerevst_code_b200.weights import synthetic_state_dict
    rng = np.random.RandomState(3)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    style, frames = smooth(64, 72), [smooth(48, 64) for _ in range(5)]
    fw = Stylization(synthetic_state_dict(0), cuda=True)
    fw.prepare_style(style)
    if backend == "gloo":      # one GPU: gather through host memory
        import rerevst_code_b200.dist as D
        orig = D.allgather_parts
        D.allgather_parts = lambda part, group=None: orig(part.cpu(), group).to(part.device)
    sharded_prepass(fw, frames, rank, world)
    st = fw.model._eng().export_clip_state()
    q.put((rank, {k: v.cpu().numpy() for k, v in st["stats"].items()},
           {k: (a.cpu().numpy(), b.cpu().numpy()) for k, (a, b) in st["filters"].items()}))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_prepass_equals_single_process():
    """world_size 2 (two GPUs if present, else both ranks on GPU 0 with a gloo gather): the sharded pre-pass
    gives the tables of the single-process pre-pass, identically on both ranks."""
    from rerevst_code_b200.framework import Stylization
    from rerevst_code_b200.weights import synthetic_state_dict
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    rng = np.random.RandomState(3)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    style, frames = smooth(64, 72), [smooth(48, 64) for _ in range(5)]
    fw = Stylization(synthetic_state_dict(0), cuda=True)
    fw.prepare_style(style)
    fw.clean()
    for f in frames:
        fw.add(f)
    fw.compute()
    ref = fw.model._eng().export_clip_state()
    from conftest import rel_linf
    for rank, stats, filters in got:
        for k, v in ref["stats"].items():
            for r in range(4):
                assert rel_linf(stats[k][r], v[r].cpu().numpy()) < 1e-4, (rank, k, r)
        for k, (a, b) in ref["filters"].items():
            assert rel_linf(filters[k][0], a.cpu().numpy()) < 1e-4 and rel_linf(filters[k][1], b.cpu().numpy()) < 1e-4
    for k in got[0][1]:
        assert np.array_equal(got[0][1][k], got[1][1][k]), k
