"""Per-layer operand-precision sensitivity of the per-frame path, simulated on the CPU oracle.
TEST INFRASTRUCTURE ONLY (runs where torch-CPU runs; no GPU).

    python -m oracle.precision_sweep [--case cfg1_256] [--modes x2a,x2w,h2w,bf16] [--cumulative]

The CUDA path carries every fp32 operand as hi = bf16(v), lo = bf16(v - hi) and issues three MMAs per
k-slice: hi*Whi + hi*Wlo + lo*Whi ("x3").  The question behind VERDICT r1 item 5: which layers could
issue fewer MMAs and keep the final frame within 5e-4 relative L-inf of the fp32 reference?
The candidates, simulated here by rounding the operands of ONE convolution (all others at x3) and
running the rest of the network in fp32:

  x3    A16 * W16            (both operands hi+lo, 16 significant bits)            3 MMAs  -- the shipped mode
  x2a   A16 * bf16(W)        (drop hi*Wlo: weights at 8 bits)                      2 MMAs
  x2w   bf16(A) * W16        (drop lo*Whi: activations at 8 bits)                  2 MMAs
  h2w   A22 * fp16(W)        (fp16 hi/lo planes, drop hi*Wlo: weights at 11 bits)  2 MMAs
  h2a   fp16(A) * W22        (fp16 planes, drop lo*Whi)                            2 MMAs
  bf16  bf16(A) * bf16(W)                                                          1 MMA

Output: final relative L-inf per (layer, mode); with --cumulative, the cheapest greedy assignment
that stays under the budget.  The numbers recorded in DESIGN.md section 4.3 come from this script.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, stylenet  # noqa: E402
from rerevst_code_b200.weights import synthetic_state_dict  # noqa: E402


def _bf16(v):
    return v.to(torch.bfloat16).float()


def _fp16(v):
    return v.to(torch.float16).float()


def _two(v, f):
    h = f(v)
    return h + f(v - h)


ROUND = {
    "x3": (lambda a: _two(a, _bf16), lambda w: _two(w, _bf16)),
    "x2a": (lambda a: _two(a, _bf16), _bf16),
    "x2w": (_bf16, lambda w: _two(w, _bf16)),
    "h2w": (lambda a: _two(a, _fp16), _fp16),
    "h2a": (_fp16, lambda w: _two(w, _fp16)),
    "bf16": (_bf16, _bf16),
    "f32": (lambda a: a, lambda w: w),
}

# the 31 convolutions of TransformerNet.forward in call order (the 6 dynamic 1x1 filters are folded into their
# neighbours on the GPU and stay exact here)
LAYERS = (["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4", "conv4_1"]
          + [f"Filter{i}.{d}" for i in (1, 2, 3) for d in ("down", "up")]
          + [f"slice{i}.{c}" for i in (4, 3, 2) for c in ("shortcut", "conv1", "conv2")] + ["slice1"])


class Patched:
    """Replaces F.conv2d inside oracle.stylenet by a version that rounds the operands of chosen calls."""

    def __init__(self, assign):
        self.assign = assign        # layer name -> mode
        self.i = 0

    def __call__(self, x, w, b=None, padding=0):
        if w.shape[-1] == 1 and w.shape[0] == 32 and w.shape[1] == 32:          # dynamic 32x32 filter
            return self.orig(x, w, b, padding=padding)
        name = LAYERS[self.i]
        self.i += 1
        fa, fw = ROUND[self.assign.get(name, "x3")]
        return self.orig(fa(x), fw(w), b, padding=padding)

    def __enter__(self):
        self.orig = F.conv2d
        stylenet.F.conv2d = self          # stylenet uses the module attribute F.conv2d
        return self

    def __exit__(self, *a):
        stylenet.F.conv2d = self.orig


def run(o, frame, assign):
    with Patched(assign) as p:
        out = o.forward(frame)
        assert p.i == len(LAYERS), p.i
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="cfg1_256")
    ap.add_argument("--modes", default="x2a,x2w,h2w,h2a,bf16")
    ap.add_argument("--budget", type=float, default=5e-4)
    ap.add_argument("--cumulative", action="store_true")
    ap.add_argument("--layers", default="", help="comma-separated prefixes of the layers to sweep (default: all)")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic_state_dict(cases.WEIGHT_SEED)
    if args.case == "full_1080p":          # the bench's frame size: 1216 x 2048 (1080p padded), pre-pass on two 270 x 480 samples
        g = cases.gen()
        style = torch.randn(1, 3, 256, 256, generator=g)
        samples = [torch.randn(1, 3, 270, 480, generator=g) for _ in range(2)]
        frame = torch.randn(1, 3, 1216, 2048, generator=g)
    else:
        style, samples, frame = cases.global_inputs(args.case)
    o = stylenet.GlobalOracle(sd)
    o.generate_style_features(style)
    o.clean()
    for s in samples:
        o.add(s)
    o.compute()
    ref = run(o, frame, {k: "f32" for k in LAYERS})
    scale = float(ref.abs().max())
    err = lambda y: float((y - ref).abs().max()) / scale
    base = err(run(o, frame, {}))
    print(f"case {args.case}: all layers x3 -> rel L-inf {base:.3e}")
    modes = args.modes.split(",")
    table = {}
    print("layer".ljust(18) + "".join(m.rjust(11) for m in modes))
    sweep = [n for n in LAYERS if not args.layers or any(n.startswith(p) for p in args.layers.split(","))]
    for name in sweep:
        row = []
        for m in modes:
            e = err(run(o, frame, {name: m}))
            table[(name, m)] = e
            row.append(e)
        print(name.ljust(18) + "".join(f"{e:11.2e}" for e in row), flush=True)
    for m in modes:
        e = err(run(o, frame, {k: m for k in sweep}))
        print(f"all swept layers {m}: {e:.3e}")
    if args.cumulative:
        for m in modes:
            if m == "bf16":
                continue
            order = sorted(sweep, key=lambda k: table[(k, m)])
            assign, kept = {}, []
            for name in order:
                trial = dict(assign)
                trial[name] = m
                e = err(run(o, frame, trial))
                if e <= args.budget:
                    assign = trial
                    kept.append((name, e))
            print(f"greedy {m} under {args.budget:g}: {[k for k, _ in kept]} -> {kept[-1][1] if kept else base:.3e}")


if __name__ == "__main__":
    main()
