"""Seeded input definitions shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE ONLY.

Inputs are regenerated from seeds; only reference OUTPUTS are stored under tests/golden/.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

WEIGHT_SEED = 0
INPUT_SEED = 1234

# name -> (style HxW, sample-frame HxW (unpadded, pre-pass), n_samples, forward-frame HxW)
GLOBAL_CASES = {
    # one 256x256 frame + one style, stats from two further 256x256 frames
    "cfg1_256": ((256, 256), (256, 256), 2, (256, 256)),
    # sample frames smaller than the (padded) frame they are applied to (quirk Q3), ragged sizes
    "small_q3": ((72, 88), (40, 56), 3, (64, 96)),
    # single sample: batch statistics degenerate to instance statistics
    "n1": ((64, 64), (48, 48), 1, (48, 48)),
}

FRAME_CASES = {
    "frame_small": ((64, 80), (56, 72)),   # style HxW, frame HxW
}

WARP_SIZES = ((256, 256), (512, 512), (436, 1024), (1080, 1920))
WARP_SMALL = ((8, 8), (7, 5), (1, 1), (1, 9), (33, 17))


def gen():
    return torch.Generator().manual_seed(INPUT_SEED)


def global_inputs(name):
    (sh, sw), (ph, pw), n, (fh, fw) = GLOBAL_CASES[name]
    g = gen()
    style = torch.randn(1, 3, sh, sw, generator=g)
    samples = [torch.randn(1, 3, ph, pw, generator=g) for _ in range(n)]
    frame = torch.randn(1, 3, fh, fw, generator=g)
    return style, samples, frame


def frame_inputs(name):
    (sh, sw), (fh, fw) = FRAME_CASES[name]
    g = gen()
    return torch.randn(1, 3, sh, sw, generator=g), torch.randn(1, 3, fh, fw, generator=g)


def warp_flow(h, w, b=2, seed=7):
    """Random flow (sigma 8 px) with one batch item forced onto rounding ties."""
    g = torch.Generator().manual_seed(seed + 131 * h + w)
    flo = torch.randn(b, 2, h, w, generator=g) * 8.0
    if b > 1:
        # x - u and y - v on exact .5 grid positions before normalisation -> near-tie after it
        flo[1, 0] = torch.arange(w).view(1, w).float() - (torch.randint(0, max(w, 1), (h, w), generator=g).float() + 0.5)
        flo[1, 1] = torch.arange(h).view(h, 1).float() - (torch.randint(0, max(h, 1), (h, w), generator=g).float() - 0.5)
    return flo.numpy()


def digest(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
