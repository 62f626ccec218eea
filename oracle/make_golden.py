"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules from
/root/reference on CPU.  TEST INFRASTRUCTURE ONLY.  Run here (the reference is not on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Weights: the package's deterministic synthetic checkpoint, loaded strictly into the reference
``TransformerNet`` (both shipped checkpoints are empty placeholders).  Inputs: oracle/cases.py.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import cases  # noqa: E402
from rerevst_code_b200.weights import synthetic_state_dict  # noqa: E402

NORM_ATTR = {
    "norm0": ("norm", 0), "norm1": ("norm", 1), "norm2": ("norm", 2), "norm3": ("norm", 3), "norm4": ("norm", 4),
}


def import_reference(module: str, sub: str):
    """Import ``module`` from /root/reference/<sub> (needs two stubs for the frame/train models:
    a dummy kornia and a vgg19 that ignores pretrained=True)."""
    if "kornia" not in sys.modules:
        k = types.ModuleType("kornia")
        k.filters = types.ModuleType("kornia.filters")
        k.filters.GaussianBlur2d = lambda *a, **kw: None
        sys.modules["kornia"] = k
        sys.modules["kornia.filters"] = k.filters
    import torchvision.models as tvm
    if not getattr(tvm.vgg19, "_rrv_patched", False):
        orig = tvm.vgg19

        def vgg19(pretrained=False, **kw):
            return orig(weights=None)
        vgg19._rrv_patched = True
        tvm.vgg19 = vgg19
    p = os.path.join(REF, sub)
    sys.path.insert(0, p)
    try:
        sys.modules.pop(module, None)
        return importlib.import_module(module)
    finally:
        sys.path.remove(p)


def reference_stats(net):
    d = net.Decoder
    mods = {"norm0": d.norm[0], "norm1": d.norm[1], "norm2": d.norm[2], "norm3": d.norm[3], "norm4": d.norm[4]}
    for b in ("slice4", "slice3", "slice2"):
        mods[b + ".norm1"] = getattr(d, b).norm1
        mods[b + ".norm2"] = getattr(d, b).norm2
    out = {}
    for k, m in mods.items():
        # after a forward the tables are expanded views; take one spatial position
        tabs = [t[:1, :, :1, :1].reshape(-1) for t in (m.saved_mean, m.saved_std, m.x_min, m.x_max)]
        out[k] = torch.stack(tabs).numpy()
    return out


def run_global(name, sd):
    g = import_reference("style_network_global", "test")
    net = g.TransformerNet().eval()
    net.load_state_dict(sd, strict=True)
    style, samples, frame = cases.global_inputs(name)
    res = {}
    with torch.no_grad():
        net.generate_style_features(style)
        net.clean()
        for s in samples:
            net.add(s)
        net.compute()
        for k, v in reference_stats(net).items():
            res["stat/" + k] = v
        for f in ("Filter1", "Filter2", "Filter3"):
            kf = getattr(net.Decoder, f)
            res[f"filter/{f}.F1"] = kf.F1.filter.reshape(32, 32).numpy()
            res[f"filter/{f}.F2"] = kf.F2.filter.reshape(32, 32).numpy()
        fs = net.F_style
        for lvl in ("relu1_1", "relu2_1", "relu3_1", "relu4_1"):
            ms = getattr(fs, lvl)
            res[f"style/{lvl}"] = torch.stack([ms.mean.reshape(-1), ms.std.reshape(-1)]).numpy()
        res["style/map"] = fs.map.numpy()
        res["F_content"] = net.Encoder(net.RGB2Gray(frame)).numpy()
        res["out"] = net(frame).numpy()
    return res


def run_frame(name, sd):
    m = import_reference("style_network_frame", "test")
    net = m.TransformerNet().eval()
    net.load_state_dict(sd, strict=True)
    style, frame = cases.frame_inputs(name)
    with torch.no_grad():
        net.generate_style_features(style)
        return {"out": net(frame).numpy()}


def run_warp():
    ln = import_reference("loss_networks", "train")
    res, meta = {}, {}
    for (h, w) in cases.WARP_SIZES + cases.WARP_SMALL:
        flo = cases.warp_flow(h, w)
        img = np.zeros((flo.shape[0], 2, h, w), np.float32)
        img[:, 0] = np.arange(w, dtype=np.float32)[None, None, :]
        img[:, 1] = np.arange(h, dtype=np.float32)[None, :, None]
        out = ln.warp(torch.from_numpy(img), torch.from_numpy(flo)).numpy()
        ix, iy = out[:, 0].astype(np.int32), out[:, 1].astype(np.int32)
        meta[f"{h}x{w}"] = {"ix_sha256": cases.digest(ix), "iy_sha256": cases.digest(iy)}
        if (h, w) in cases.WARP_SMALL:
            res[f"ix/{h}x{w}"], res[f"iy/{h}x{w}"] = ix, iy
    # TemporalLoss.forward on a 64x64 pair
    g = torch.Generator().manual_seed(99)
    first = torch.randn(2, 3, 64, 64, generator=g)
    second = torch.randn(2, 3, 64, 64, generator=g)
    flo = torch.from_numpy(cases.warp_flow(64, 64))
    loss, warped = ln.TemporalLoss().forward(first, second, flo)
    res["tl/loss"] = np.array(float(loss), np.float64)
    res["tl/warped"] = warped.numpy()
    return res, meta


def run_train(sd):
    """train/style_networks.py: validation (frame mode without RGB2Gray, :556-559), the Vgg19 loss network (:284-314) and
    style_loss / content_loss (:503-516) on the temporal-loss path of train/train.py:375-388."""
    m = import_reference("style_networks", "train")
    net = m.TransformerNet().eval()
    net.load_state_dict(sd, strict=True)
    style, frame = cases.frame_inputs("frame_small")
    g = torch.Generator().manual_seed(4321)
    other = torch.randn(2, 3, 40, 56, generator=g)
    with torch.no_grad():
        out = net.validation(frame, style)
        f_a = net.Vgg19(other)
        f_b = net.Vgg19(torch.flip(other, dims=(0, 3)))
        res = {"validation": out.numpy(), "style_loss": np.array(float(net.style_loss(f_a, f_b)), np.float64),
               "content_loss": np.array(float(net.content_loss(f_a, f_b)), np.float64), "relu4_1": f_a.relu4_1.numpy()}
        for lvl, ft in zip(f_a._fields, f_a):
            mean, std = m.calc_mean_std(ft)
            res[f"mean/{lvl}"], res[f"std/{lvl}"] = mean.numpy(), std.numpy()
    return res


def multi_inputs():
    g = torch.Generator().manual_seed(777)
    styles = [torch.randn(1, 3, 64, 72, generator=g), torch.randn(1, 3, 56, 64, generator=g)]
    patches = [torch.randn(1, 3, 48, 64, generator=g) for _ in range(2)]
    frame = torch.randn(1, 3, 64, 80, generator=g)
    return styles, patches, frame


def run_multi(sd):
    """"Multi-style Interpolation/style_network.py": two styles, statistics from two patches, three weightings."""
    m = import_reference("style_network", "Multi-style Interpolation")
    net = m.TransformerNet(style_num=2).eval()
    net.load_state_dict(sd, strict=True)
    styles, patches, frame = multi_inputs()
    res = {}
    with torch.no_grad():
        for i, s in enumerate(styles):
            net.generate_style_features(s, i)
        net.clean()
        for pch in patches:
            net.add_patch(net.generate_content_features(pch))
        net.compute_norm()
        fc = net.generate_content_features(frame)
        for name, w in (("w10", [1.0, 0.0]), ("w37", [0.3, 0.7]), ("w55", [0.5, 0.5])):
            res["out/" + name] = net(fc, w).numpy()
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    sd = synthetic_state_dict(cases.WEIGHT_SEED)
    for name in cases.GLOBAL_CASES:
        res = run_global(name, sd)
        if name == "cfg1_256":      # keep the fixture small: drop the two big maps
            res.pop("style/map"); res.pop("F_content")
        np.savez_compressed(os.path.join(OUT, f"global_{name}.npz"), **res)
        print(name, {k: v.shape for k, v in res.items() if not k.startswith("stat")})
    for name in cases.FRAME_CASES:
        np.savez_compressed(os.path.join(OUT, f"frame_{name}.npz"), **run_frame(name, sd))
    np.savez_compressed(os.path.join(OUT, "train_model.npz"), **run_train(sd))
    np.savez_compressed(os.path.join(OUT, "multi_style.npz"), **run_multi(sd))
    res, meta = run_warp()
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **res)
    with open(os.path.join(OUT, "warp_digests.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("Golden outputs of the unmodified reference (commit b7f39f2) on CPU, torch %s.\n"
                "Generated by `python -m oracle.make_golden`; inputs are re-derived from seeds in oracle/cases.py.\n"
                % torch.__version__)


if __name__ == "__main__":
    main()
