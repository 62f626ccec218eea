"""CPU: the oracle against the golden fixtures generated from the unmodified reference, and
(where /root/reference exists) against the reference modules themselves."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, HAVE_REFERENCE, rel_linf
from oracle import cases, stylenet, warp as owarp

# The oracle issues the same torch CPU ops as the reference; the residual is thread-count /
# oneDNN blocking noise (8.5e-7 between 1 and 8 threads).
ORACLE_TOL = 2e-5


def _run_global(name, sd):
    style, samples, frame = cases.global_inputs(name)
    o = stylenet.GlobalOracle(sd)
    o.generate_style_features(style)
    o.clean()
    for s in samples:
        o.add(s)
    o.compute()
    return o, o.forward(frame)


@pytest.mark.parametrize("name", list(cases.GLOBAL_CASES))
def test_global_mode_matches_golden(name, state_dict):
    gold = np.load(os.path.join(GOLDEN, f"global_{name}.npz"))
    o, out = _run_global(name, state_dict)
    assert rel_linf(out.numpy(), gold["out"]) < ORACLE_TOL
    for k in stylenet.ClipState.NAMES:
        st = o.clip.stats[k]
        tab = torch.stack([t.reshape(-1) for t in st]).numpy()
        assert tab.shape == gold["stat/" + k].shape
        for row in range(4):
            assert rel_linf(tab[row], gold["stat/" + k][row]) < ORACLE_TOL, (k, row)
    for f in ("Filter1", "Filter2", "Filter3"):
        for j, p in enumerate(("F1", "F2")):
            assert rel_linf(o.clip.filters[f][j].reshape(32, 32).numpy(), gold[f"filter/{f}.{p}"]) < ORACLE_TOL
    for lvl in ("relu1_1", "relu2_1", "relu3_1", "relu4_1"):
        ms = getattr(o.F_style, lvl)
        got = torch.stack([ms.mean.reshape(-1), ms.std.reshape(-1)]).numpy()
        assert rel_linf(got, gold[f"style/{lvl}"]) < ORACLE_TOL


@pytest.mark.parametrize("name", list(cases.FRAME_CASES))
def test_frame_mode_matches_golden(name, state_dict):
    gold = np.load(os.path.join(GOLDEN, f"frame_{name}.npz"))
    style, frame = cases.frame_inputs(name)
    fs = stylenet.encoder_style(style, state_dict)
    out = stylenet.frame_mode_forward(state_dict, frame, fs)
    assert rel_linf(out.numpy(), gold["out"]) < ORACLE_TOL


def test_train_model_matches_golden(state_dict):
    """train/style_networks.py on the temporal-loss path: validation = frame mode without RGB2Gray,
    the Vgg19 loss network, calc_mean_std, style_loss and content_loss -- against the unmodified reference's outputs."""
    gold = np.load(os.path.join(GOLDEN, "train_model.npz"))
    style, frame = cases.frame_inputs("frame_small")
    fs = stylenet.encoder_style(style, state_dict)
    out = stylenet.frame_mode_forward(state_dict, frame, fs, gray=False)
    assert rel_linf(out.numpy(), gold["validation"]) < ORACLE_TOL
    g = torch.Generator().manual_seed(4321)
    other = torch.randn(2, 3, 40, 56, generator=g)
    fa = stylenet.vgg19_features(other, state_dict)
    fb = stylenet.vgg19_features(torch.flip(other, dims=(0, 3)), state_dict)
    assert rel_linf(fa[3].numpy(), gold["relu4_1"]) < ORACLE_TOL
    sl = 0.0
    for lvl, a, b in zip(("relu1_1", "relu2_1", "relu3_1", "relu4_1"), fa, fb):
        ma, mb = stylenet.cal_mean_std(a), stylenet.cal_mean_std(b)
        assert rel_linf(ma.mean.numpy(), gold[f"mean/{lvl}"]) < ORACLE_TOL and rel_linf(ma.std.numpy(), gold[f"std/{lvl}"]) < ORACLE_TOL
        sl = sl + torch.nn.functional.mse_loss(ma.mean, mb.mean) + torch.nn.functional.mse_loss(ma.std, mb.std)
    assert abs(float(sl) - float(gold["style_loss"])) < 1e-5 * abs(float(gold["style_loss"]))
    cl = torch.nn.functional.mse_loss(fa[3], fb[3])
    assert abs(float(cl) - float(gold["content_loss"])) < 1e-5 * abs(float(gold["content_loss"]))


def test_multi_style_matches_golden(state_dict):
    """"Multi-style Interpolation/style_network.py" (SURVEY 8f N3): per-style statistics, blended by style_weight in forward."""
    from oracle.make_golden import multi_inputs
    gold = np.load(os.path.join(GOLDEN, "multi_style.npz"))
    styles, patches, frame = multi_inputs()
    o = stylenet.MultiStyleOracle(state_dict, 2)
    for i, s in enumerate(styles):
        o.generate_style_features(s, i)
    for pch in patches:
        o.add_patch(o.generate_content_features(pch))
    o.compute_norm()
    fc = o.generate_content_features(frame)
    for name, w in (("w10", [1.0, 0.0]), ("w37", [0.3, 0.7]), ("w55", [0.5, 0.5])):
        assert rel_linf(o.forward(fc, w).numpy(), gold["out/" + name]) < ORACLE_TOL


def test_q1_only_first_sample_is_filtered(state_dict):
    """Quirk Q1: in the pre-pass only sample 0 goes through the dynamic filters and
    its residual is broadcast to every sample."""
    g = torch.Generator().manual_seed(3)
    content = torch.randn(3, 512, 5, 6, generator=g)
    style = torch.randn(1, 512, 4, 4, generator=g)
    out, wf1, wf2 = stylenet.kernel_filter_compute(state_dict, "Decoder.Filter1", content, style)
    res = out - content
    assert torch.allclose(res[1], res[0], atol=1e-5) and torch.allclose(res[2], res[0], atol=1e-5)
    single = stylenet.kernel_filter(state_dict, "Decoder.Filter1", content[:1], wf1, wf2)
    assert torch.allclose(single, out[:1], atol=1e-5)


def test_work_model_matches_baseline_md():
    assert abs(stylenet.conv_flops_per_frame(256, 256) / 1e9 - 80.39) < 0.01
    assert abs(stylenet.conv_flops_per_frame(1216, 2048) / 1e9 - 3054.90) < 0.05
    assert abs(stylenet.conv_flops_per_frame(896, 1408) / 1e9 - 1547.55) < 0.05


# ---------------------------------------------------------------- warp

def _index_image(b, h, w):
    img = np.zeros((b, 2, h, w), np.float32)
    img[:, 0] = np.arange(w, dtype=np.float32)[None, None, :]
    img[:, 1] = np.arange(h, dtype=np.float32)[None, :, None]
    return img


@pytest.mark.parametrize("hw", cases.WARP_SIZES + cases.WARP_SMALL)
def test_warp_indices_match_golden(hw):
    h, w = hw
    with open(os.path.join(GOLDEN, "warp_digests.json")) as f:
        meta = json.load(f)[f"{h}x{w}"]
    iy, ix = owarp.warp_indices(cases.warp_flow(h, w))
    assert cases.digest(ix) == meta["ix_sha256"] and cases.digest(iy) == meta["iy_sha256"]
    if hw in cases.WARP_SMALL:
        gold = np.load(os.path.join(GOLDEN, "warp.npz"))
        assert np.array_equal(ix, gold[f"ix/{h}x{w}"]) and np.array_equal(iy, gold[f"iy/{h}x{w}"])


@pytest.mark.parametrize("hw", ((64, 64), (37, 91), (1, 1), (2, 300)))
def test_warp_matches_torch_grid_sample(hw):
    """Pin against the third-party op the reference calls (loss_networks.py:37)."""
    import torch.nn.functional as F
    h, w = hw
    flo = cases.warp_flow(h, w)
    b = flo.shape[0]
    xx = torch.arange(0, w).view(1, -1).repeat(h, 1).view(1, 1, h, w).repeat(b, 1, 1, 1)
    yy = torch.arange(0, h).view(-1, 1).repeat(1, w).view(1, 1, h, w).repeat(b, 1, 1, 1)
    vgrid = torch.cat((xx, yy), 1).float() - torch.from_numpy(flo)
    vgrid[:, 0] = 2.0 * vgrid[:, 0] / max(w - 1, 1) - 1.0
    vgrid[:, 1] = 2.0 * vgrid[:, 1] / max(h - 1, 1) - 1.0
    img = _index_image(b, h, w)
    ref = F.grid_sample(torch.from_numpy(img), vgrid.permute(0, 2, 3, 1), padding_mode="border",
                        mode="nearest", align_corners=False).numpy()
    iy, ix = owarp.warp_indices(flo)
    assert np.array_equal(ref[:, 0].astype(np.int32), ix) and np.array_equal(ref[:, 1].astype(np.int32), iy)
    assert np.array_equal(owarp.warp(img, flo), ref)


def test_warp_c_oracle_equals_numpy_oracle():
    import ctypes
    so = os.path.join(os.path.dirname(os.path.abspath(owarp.__file__)), "liboracle_warp.so")
    if not os.path.exists(so):
        pytest.skip("oracle/liboracle_warp.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(so)
    for (h, w) in ((33, 17), (64, 80), (1, 9)):
        flo = cases.warp_flow(h, w)
        b = flo.shape[0]
        iy = np.zeros((b, h, w), np.int32)
        ix = np.zeros((b, h, w), np.int32)
        lib.rrv_oracle_warp_indices(flo.ctypes.data_as(ctypes.c_void_p), b, h, w,
                                    iy.ctypes.data_as(ctypes.c_void_p), ix.ctypes.data_as(ctypes.c_void_p))
        ny, nx = owarp.warp_indices(flo)
        assert np.array_equal(iy, ny) and np.array_equal(ix, nx)


def test_temporal_loss_matches_golden():
    gold = np.load(os.path.join(GOLDEN, "warp.npz"))
    g = torch.Generator().manual_seed(99)
    first = torch.randn(2, 3, 64, 64, generator=g).numpy()
    second = torch.randn(2, 3, 64, 64, generator=g).numpy()
    loss, warped = owarp.temporal_loss(first, second, cases.warp_flow(64, 64))
    assert np.array_equal(warped, gold["tl/warped"])
    assert abs(loss - float(gold["tl/loss"])) <= 1e-6 * abs(float(gold["tl/loss"]))


# ---------------------------------------------------------------- live reference (this container only)

@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference not present (GPU box)")
def test_oracle_equals_live_reference(state_dict):
    from oracle.make_golden import run_global
    res = run_global("small_q3", state_dict)
    o, out = _run_global("small_q3", state_dict)
    assert rel_linf(out.numpy(), res["out"]) < ORACLE_TOL
    taps = {}
    _, _, frame = cases.global_inputs("small_q3")
    o.forward(frame, taps)
    assert rel_linf(taps["F_content"].numpy(), res["F_content"]) < ORACLE_TOL
