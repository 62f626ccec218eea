"""CPU oracle for the optical-flow warp of the temporal loss.  TEST INFRASTRUCTURE ONLY
(imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs; never by
the product package).

Restates ``train/loss_networks.py:20-38`` (``warp``) and ``:106-111``
(``TemporalLoss.forward``) in numpy float32, operation for operation:

    vgrid = grid - flo                                   (:30)
    gx    = 2.0 * vgrid_x / max(W-1, 1) - 1.0            (:33)   three separately rounded fp32 ops
    gy    = 2.0 * vgrid_y / max(H-1, 1) - 1.0            (:34)
    F.grid_sample(x, vgrid, mode='nearest', padding_mode='border')   (:37)

``align_corners`` is not passed, so torch >= 1.3 un-normalises with
``((g + 1) * size - 1) / 2`` (ATen ``grid_sampler_unnormalize``), clips to ``[0, size-1]``
(``clip_coordinates`` for border padding) and rounds with ``nearbyint`` (ties to even).
The source indices are integers; parity with the CUDA kernel is exact equality.  The
restatement is pinned against torch's own ``F.grid_sample`` and against the reference
function itself in ``tests/test_oracle_warp.py`` and by ``tests/golden/warp_*.npz``.
"""
from __future__ import annotations

import numpy as np

_f32 = np.float32


def warp_indices(flo: np.ndarray):
    """flo: [B,2,H,W] float32 (ch0 = x flow, ch1 = y flow, in pixels).
    Returns (iy, ix) int32 arrays [B,H,W]: the source pixel every output pixel gathers."""
    flo = np.asarray(flo, dtype=_f32)
    b, _, h, w = flo.shape
    xx = np.arange(w, dtype=_f32)[None, None, :]
    yy = np.arange(h, dtype=_f32)[None, :, None]
    vx = (xx - flo[:, 0]).astype(_f32)
    vy = (yy - flo[:, 1]).astype(_f32)
    gx = ((_f32(2.0) * vx).astype(_f32) / _f32(max(w - 1, 1))).astype(_f32) - _f32(1.0)
    gy = ((_f32(2.0) * vy).astype(_f32) / _f32(max(h - 1, 1))).astype(_f32) - _f32(1.0)
    gx = gx.astype(_f32)
    gy = gy.astype(_f32)
    # grid_sampler_unnormalize(align_corners=False): ((coord + 1) * size - 1) / 2
    ux = ((((gx + _f32(1.0)).astype(_f32) * _f32(w)).astype(_f32) - _f32(1.0)).astype(_f32) / _f32(2.0)).astype(_f32)
    uy = ((((gy + _f32(1.0)).astype(_f32) * _f32(h)).astype(_f32) - _f32(1.0)).astype(_f32) / _f32(2.0)).astype(_f32)
    # clip_coordinates: min(size-1, max(coord, 0))
    ux = np.minimum(_f32(w - 1), np.maximum(ux, _f32(0.0)))
    uy = np.minimum(_f32(h - 1), np.maximum(uy, _f32(0.0)))
    ix = np.rint(ux).astype(np.int32)  # nearbyint: round half to even
    iy = np.rint(uy).astype(np.int32)
    return iy, ix


def warp(x: np.ndarray, flo: np.ndarray) -> np.ndarray:
    """warp(x, flo, padding_mode='border'), loss_networks.py:20-38.  x: [B,C,H,W]."""
    x = np.asarray(x)
    iy, ix = warp_indices(flo)
    b = np.arange(x.shape[0])[:, None, None]
    return np.ascontiguousarray(np.moveaxis(x[b, :, iy, ix], -1, 1))


def temporal_loss(first: np.ndarray, second: np.ndarray, flo: np.ndarray):
    """TemporalLoss.forward, loss_networks.py:106-111: (mean |warp(first) - second|, warped).
    The mean is accumulated in float64 and is compared with a 1e-6 relative tolerance."""
    w = warp(first, flo)
    return float(np.mean(np.abs(w.astype(np.float64) - np.asarray(second, dtype=np.float64)))), w


def fake_flow(height: int, width: int, seed: int = 0, motion_level: float = 8.0, shift_level: int = 10):
    """Smooth synthetic flow with the statistics of TemporalLoss.GenerateFakeFlow
    (loss_networks.py:71-86: N(0, 8) on a /100 grid, resized, + global shift in [-10, 10],
    box-blurred).  cv2 is not needed: bilinear resize and a separable box blur in numpy."""
    rng = np.random.RandomState(seed)
    gh, gw = max(height // 100, 1) + 1, max(width // 100, 1) + 1
    coarse = rng.normal(0, motion_level, size=(gh, gw, 2))
    ys = np.linspace(0, gh - 1, height)
    xs = np.linspace(0, gw - 1, width)
    y0 = np.clip(np.floor(ys).astype(int), 0, gh - 2)
    x0 = np.clip(np.floor(xs).astype(int), 0, gw - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    c = coarse
    flow = ((1 - fy) * (1 - fx) * c[y0][:, x0] + (1 - fy) * fx * c[y0][:, x0 + 1]
            + fy * (1 - fx) * c[y0 + 1][:, x0] + fy * fx * c[y0 + 1][:, x0 + 1])
    flow[:, :, 0] += rng.randint(-shift_level, shift_level + 1)
    flow[:, :, 1] += rng.randint(-shift_level, shift_level + 1)
    k = min(25, height // 4, width // 4)
    if k > 1:
        ker = np.ones(k) / k
        flow = np.apply_along_axis(lambda v: np.convolve(v, ker, mode="same"), 0, flow)
        flow = np.apply_along_axis(lambda v: np.convolve(v, ker, mode="same"), 1, flow)
    return np.ascontiguousarray(flow.transpose(2, 0, 1)).astype(np.float32)
