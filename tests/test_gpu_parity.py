"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures of the unmodified reference.  Tolerances follow BASELINE.json: <= 1e-3 relative L-inf
for the fp32-accurate ("x3") path, exact equality for the warp's integer indices."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_linf

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: 1e-3 relative L-inf (fp32)
TIGHT = 2e-5        # single fp32 layers (FFMA path: operands carry >= 16 mantissa bits)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def L():
    from rerevst_code_b200 import _lib
    _lib.lib()
    return _lib


def _impls(L):
    out = [("ffma", L.IMPL_FFMA)]
    if L.lib().rrv_tc_weight_bytes(64, 64, 3, 0) > 0:
        out.append(("tc", L.IMPL_TCGEN05))
    return out


def _to_planes(L, x_nchw, x3=True):
    from rerevst_code_b200.engine import Planes
    N, Cc, H, W = x_nchw.shape
    p = Planes(N, H, W, Cc, x3, x_nchw.device)
    xc = x_nchw.contiguous()
    L.check(L.lib().rrv_nchw_to_planes(xc.data_ptr(), N, H, W, Cc, L.ptr(p.hi), L.ptr(p.lo), L.stream()))
    return p


def _from_planes(L, p):
    out = torch.empty((p.N, p.C, p.H, p.W), dtype=torch.float32, device=p.hi.device)
    L.check(L.lib().rrv_planes_to_nchw(L.ptr(p.hi), L.ptr(p.lo), p.N, p.H, p.W, p.C, out.data_ptr(), L.stream()))
    return out


def test_native_library_is_loaded_in_tree(L):
    maps = open("/proc/self/maps").read()
    assert "rerevst-code_b200/csrc/librerevst_b200.so" in maps


def test_planes_roundtrip_keeps_16_bits(L, dev):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 16, 9, 11, generator=g)
    for fmt, bits in ((0, 2.0 ** -16), (1, 2.0 ** -18)):
        L.check(L.lib().rrv_set_lo_format(fmt))
        y = _from_planes(L, _to_planes(L, x.to(dev))).cpu()
        assert float(((y - x).abs() / x.abs().clamp_min(0.05)).max()) < bits
    L.check(L.lib().rrv_set_lo_format(0))


# ---------------------------------------------------------------------------------- single layers

CONV_CASES = [
    # N, H, W, Cin, Cout, k, ups
    (1, 16, 16, 64, 64, 3, 0),
    (2, 24, 40, 64, 128, 3, 0),
    (1, 13, 19, 128, 64, 3, 0),          # ragged: not a multiple of any tile
    (1, 16, 32, 256, 128, 3, 1),         # nearest x2 folded into the gather
    (1, 12, 20, 512, 256, 1, 0),         # 1x1 shortcut
    (1, 9, 150, 64, 64, 3, 0),           # wider than one row tile
    (1, 40, 8, 128, 512, 3, 0),
    (1, 24, 48, 256, 512, 3, 0),         # two Cout tiles of 256
    (2, 18, 22, 512, 64, 3, 0),          # KernelFilter.down_sample (padded to 64)
    (1, 26, 38, 128, 64, 3, 1),          # ups, ragged low-res size (13 x 19)
    (3, 8, 8, 64, 64, 1, 0),
    (2, 21, 37, 32, 128, 3, 0),          # 32-channel input: 64-byte operand rows (SWIZZLE_64B), ragged
    (1, 40, 24, 32, 512, 3, 0),          # KernelFilter.upsample without channel padding
    (1, 10, 12, 32, 64, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_matches_torch_cpu(L, dev, case):
    from rerevst_code_b200.engine import ConvW, make_epilogue
    N, H, W, Cin, Cout, k, ups = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    hin, win = (H // 2, W // 2) if ups else (H, W)
    x = torch.randn(N, Cin, hin, win, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    ref = F.leaky_relu(F.conv2d(xin, w, b, padding=k // 2), 0.2)
    cw = ConvW(w.to(dev), b.to(dev), ups=bool(ups))
    xp = _to_planes(L, x.to(dev))
    xq = _from_planes(L, xp).cpu()                     # what the kernel really sees (16+ bit operands)
    xin_q = F.interpolate(xq, scale_factor=2, mode="nearest") if ups else xq
    ref_q = F.leaky_relu(F.conv2d(xin_q, w, b, padding=k // 2), 0.2)
    for name, impl in _impls(L):
        d = L.Conv()
        d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, k, ups
        d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
        d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
        d.ep = make_epilogue(bias=cw.bias, act=2)
        d.out_mode = L.OUT_F32_NHWC
        out = torch.empty((N, H, W, Cout), dtype=torch.float32, device=dev)
        d.out_f32 = out.data_ptr()
        L.check(L.lib().rrv_conv2d(C.byref(d), impl, L.stream()), name)
        got = out.permute(0, 3, 1, 2).cpu()
        assert rel_linf(got, ref_q) < (TIGHT if name == "ffma" else 2e-4), name
        assert rel_linf(got, ref) < 2e-4, name


@pytest.mark.parametrize("Cout", [3, 20])
def test_conv_rgb_head_nchw(L, dev, Cout):
    """Decoder.slice1 (64 -> 3): fp32 NCHW output, Cout not a multiple of 8."""
    from rerevst_code_b200.engine import ConvW, make_epilogue
    g = torch.Generator().manual_seed(21)
    N, H, W, Cin = 2, 21, 35, 64
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 24.0
    b = torch.randn(Cout, generator=g) * 0.1
    cw = ConvW(w.to(dev), b.to(dev))
    xp = _to_planes(L, x.to(dev))
    ref = F.conv2d(_from_planes(L, xp).cpu(), w, b, padding=1)
    for name, impl in _impls(L):
        if name == "ffma" and Cout >= 8:
            continue
        d = L.Conv()
        d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, 3, 0
        d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
        d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
        d.ep = make_epilogue(bias=cw.bias)
        out = torch.full((N, 3, H, W), float("nan"), dtype=torch.float32, device=dev)
        d.out_mode, d.out_f32, d.out_C = L.OUT_F32_NCHW, out.data_ptr(), 3
        L.check(L.lib().rrv_conv2d(C.byref(d), impl, L.stream()), name)
        assert rel_linf(out.cpu(), ref[:, :3]) < 2e-4, name


def test_conv_bf16_mode(L, dev):
    """BASELINE config 3: bf16 operands (hi planes only), fp32 accumulate.  Exact up to fp32 summation
    order against a reference fed the same bf16-rounded operands."""
    if len(_impls(L)) < 2:
        pytest.skip("tcgen05 path not built")
    from rerevst_code_b200.engine import ConvW, Planes, make_epilogue
    g = torch.Generator().manual_seed(31)
    N, H, W, Cin, Cout = 1, 24, 40, 128, 128
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 34.0
    b = torch.randn(Cout, generator=g) * 0.1
    cw = ConvW(w.to(dev), b.to(dev))
    xp = _to_planes(L, x.to(dev), x3=False)
    xq = _from_planes(L, xp).cpu()
    assert torch.equal(xq, x.bfloat16().float())
    ref = F.relu(F.conv2d(xq, w.bfloat16().float(), b, padding=1))
    d = L.Conv()
    d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, 3, 0
    d.in_hi, d.in_lo = L.ptr(xp.hi), 0
    d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
    d.ep = make_epilogue(bias=cw.bias, act=1)
    o = Planes(N, H, W, Cout, False, dev)
    d.out_mode, d.out_hi, d.out_lo = L.OUT_PLANES, L.ptr(o.hi), 0
    L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()))
    got = _from_planes(L, o).cpu()
    assert rel_linf(got, ref) < 2.0 ** -8          # output rounded to bf16
    assert rel_linf(got, ref.bfloat16().float()) < 2.0 ** -7


@pytest.mark.parametrize("size", [(2, 16, 24), (2, 38, 70), (1, 52, 126), (3, 8, 16)])
def test_conv_full_epilogue_chain(L, dev, size):
    """bias -> LeakyReLU -> saved-stat norm -> + half-res residual -> saved-stat norm -> AdaIN (ResidualBlock.conv2 of slice2:
    merged-tap main loop, residual tile by TMA, output through the staging rows and TMA stores; ragged tile edges, batches)."""
    from rerevst_code_b200.engine import ConvW, make_epilogue
    from oracle import stylenet
    g = torch.Generator().manual_seed(11)
    (N, H, W), Cc = size, 64
    x = torch.randn(N, Cc, H, W, generator=g)
    w = torch.randn(Cc, Cc, 3, 3, generator=g) / 24.0
    b = torch.randn(Cc, generator=g) * 0.1
    res = torch.randn(N, Cc, H // 2, W // 2, generator=g)
    y = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.2)
    st1, _ = stylenet.in_compute(y)
    # clamp tighter than the data so the max/min stages are exercised
    st1 = stylenet.SavedStat(st1.mean, st1.rstd, st1.lo * 0.5, st1.hi * 0.5)
    y1 = stylenet.in_forward(y, st1) + F.interpolate(res, scale_factor=2, mode="nearest")
    st2, _ = stylenet.in_compute(y1)
    st2 = stylenet.SavedStat(st2.mean, st2.rstd, st2.lo * 0.7, st2.hi * 0.7)
    sc, sh = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g)
    ref = stylenet.in_forward(y1, st2) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    tab = lambda s: torch.stack([t.reshape(-1) for t in s]).contiguous().to(dev)
    cw = ConvW(w.to(dev), b.to(dev))
    xp, rp = _to_planes(L, x.to(dev)), _to_planes(L, res.to(dev))
    t1, t2, aff = tab(st1), tab(st2), torch.stack([sc, sh]).contiguous().to(dev)
    rf = res.permute(0, 2, 3, 1).contiguous().to(dev)            # the same residual as an fp32 NHWC tensor (rrv_epilogue.res_f32)
    for name, impl in _impls(L):
        for residual in (rp, rf):
            d = L.Conv()
            d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cc, Cc, 3, 0
            d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
            d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
            d.ep = make_epilogue(bias=cw.bias, act=2, norm1=t1, res=residual, res_shift=1, norm2=t2, affine=aff)
            from rerevst_code_b200.engine import Planes
            o = Planes(N, H, W, Cc, True, dev)
            d.out_mode, d.out_hi, d.out_lo = L.OUT_PLANES, L.ptr(o.hi), L.ptr(o.lo)
            L.check(L.lib().rrv_conv2d(C.byref(d), impl, L.stream()), name)
            assert rel_linf(_from_planes(L, o).cpu(), ref) < 3e-4, (name, residual is rf)


@pytest.mark.parametrize("size", [(1, 19, 40), (2, 36, 21), (1, 152, 64)])
@pytest.mark.parametrize("cin,chain,pair", [(32, False, 1), (32, True, 1), (64, True, 1), (32, True, 0), (64, False, 0)])
def test_kernelfilter_upsample_epilogue(L, dev, size, cin, chain, pair):
    """KernelFilter.forward's `x + upsample(t)` (style_network_global.py:217), for Filter3 followed by Decoder.norm[1] + AdaIN
    (:443): conv3x3 cin -> 512 + bias + residual planes at the same resolution [+ saved-stat norm + affine].  On the tensor-core
    path this is the staged row-reuse epilogue (residual by cp.async one chunk ahead, chain in place, TMA store) over 64-byte
    (cin = 32) or 128-byte operand rows; ragged tile edges, a batch, more than one tile per worker."""
    from rerevst_code_b200.engine import ConvW, Planes, make_epilogue
    from oracle import stylenet
    g = torch.Generator().manual_seed(17)
    (N, H, W), Cc = size, 512
    t = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(Cc, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.randn(Cc, generator=g) * 0.1
    res = torch.randn(N, Cc, H, W, generator=g)
    cw = ConvW(w.to(dev), b.to(dev))
    tp, rp = _to_planes(L, t.to(dev)), _to_planes(L, res.to(dev))
    ref = F.conv2d(_from_planes(L, tp).cpu(), w, b, padding=1) + _from_planes(L, rp).cpu()
    kw = {}
    if chain:
        st, _ = stylenet.in_compute(ref)
        st = stylenet.SavedStat(st.mean, st.rstd, st.lo * 0.6, st.hi * 0.6)
        sc, sh = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g)
        ref = stylenet.in_forward(ref, st) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        kw = dict(norm2=torch.stack([x.reshape(-1) for x in st]).contiguous().to(dev), affine=torch.stack([sc, sh]).contiguous().to(dev))
    L.check(L.lib().rrv_tc_tune_pair(pair, 64))            # pair = 0: the single-CTA instantiation of the same path
    try:
        for name, impl in _impls(L):
            d = L.Conv()
            d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, cin, Cc, 3, 0
            d.in_hi, d.in_lo = L.ptr(tp.hi), L.ptr(tp.lo)
            d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
            d.ep = make_epilogue(bias=cw.bias, res=rp, **kw)
            o = Planes(N, H, W, Cc, True, dev)
            o.hi.fill_(float("nan"))
            d.out_mode, d.out_hi, d.out_lo = L.OUT_PLANES, L.ptr(o.hi), L.ptr(o.lo)
            L.check(L.lib().rrv_conv2d(C.byref(d), impl, L.stream()), name)
            assert rel_linf(_from_planes(L, o).cpu(), ref) < 2e-4, (name, size, cin, chain, pair)
    finally:
        L.check(L.lib().rrv_tc_tune_pair(1, 64))


@pytest.mark.parametrize("kind,gray", [(0, 1), (0, 0), (1, 1), (1, 0)])
def test_first_layer_matches_oracle(L, dev, state_dict, kind, gray):
    from oracle import stylenet
    g = torch.Generator().manual_seed(5)
    H, W = 37, 53
    w, b = state_dict["Encoder.slice.0.weight"], state_dict["Encoder.slice.0.bias"]
    if kind == 0:
        x = torch.randn(2, 3, H, W, generator=g)
        src = x.to(dev)
    else:
        u8 = torch.randint(0, 256, (2, H, W, 3), generator=g, dtype=torch.uint8)
        x = torch.cat([stylenet.transform_image(stylenet.numpy2tensor(u8[i].numpy())) for i in range(2)], 0)
        src = u8.to(dev)
    xin = stylenet.rgb2gray(x) if gray else x
    ref = F.relu(F.conv2d(xin, w, b, padding=1))
    out = torch.empty((2, H, W, 64), dtype=torch.float32, device=dev)
    wd, bd = w.to(dev), b.to(dev)                      # keep alive: data_ptr() of a temporary dangles
    L.check(L.lib().rrv_first_layer(src.data_ptr(), kind, gray, 2, H, W, wd.data_ptr(), bd.data_ptr(),
                                    0, 0, out.data_ptr(), L.stream()))
    assert rel_linf(out.permute(0, 3, 1, 2).cpu(), ref) < TIGHT


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("shape", [(2, 37, 52), (1, 8, 32), (1, 3, 100), (3, 40, 64), (1, 1, 4), (1, 9, 36), (2, 37, 53), (1, 300, 2048)])
@pytest.mark.parametrize("with_lo", [True, False])
def test_first_layer_gray_planes_kernel(L, dev, state_dict, kind, shape, with_lo):
    """The per-frame instantiation (uint8 frame, gray, planes out, W % 4 == 0: first_layer_gray_rows_kernel, one image row x 32
    columns per warp) against the fp32-output kernel and the oracle on ragged sizes: partial column tiles, border rows / columns,
    fewer units than warps, more units than warps (300 x 2048); the other shapes / sources take the tile kernel."""
    from oracle import stylenet
    from rerevst_code_b200.engine import Planes
    g = torch.Generator().manual_seed(7)
    N, H, W = shape
    w, b = state_dict["Encoder.slice.0.weight"], state_dict["Encoder.slice.0.bias"]
    if kind == 0:
        x = torch.randn(N, 3, H, W, generator=g)
        src = x.to(dev)
    else:
        u8 = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8)
        x = torch.cat([stylenet.transform_image(stylenet.numpy2tensor(u8[i].numpy())) for i in range(N)], 0)
        src = u8.to(dev)
    ref = F.relu(F.conv2d(stylenet.rgb2gray(x), w, b, padding=1))
    wd, bd = w.to(dev), b.to(dev)
    f32 = torch.empty((N, H, W, 64), dtype=torch.float32, device=dev)
    L.check(L.lib().rrv_first_layer(src.data_ptr(), kind, 1, N, H, W, wd.data_ptr(), bd.data_ptr(), 0, 0, f32.data_ptr(), L.stream()))
    o = Planes(N, H, W, 64, with_lo, dev)
    o.hi.fill_(float("nan"))                            # every element must be written
    if with_lo:
        o.lo.fill_(0x7fc0)
    L.check(L.lib().rrv_first_layer(src.data_ptr(), kind, 1, N, H, W, wd.data_ptr(), bd.data_ptr(), L.ptr(o.hi), L.ptr(o.lo), 0, L.stream()))
    got = _from_planes(L, o)
    assert torch.isfinite(got).all()
    tol = 2.0 ** -16 if with_lo else 2.0 ** -8
    scale = float(f32.abs().max()) + 1e-30
    assert float((got.permute(0, 2, 3, 1) - f32).abs().max()) / scale < tol
    assert rel_linf(got.cpu(), ref) < (TIGHT if with_lo else 5e-3)


def test_maxpool_matches_torch(L, dev):
    from rerevst_code_b200.engine import Planes
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 13, 18, generator=g)       # odd height: floor like nn.MaxPool2d
    xp = _to_planes(L, x.to(dev))
    xq = _from_planes(L, xp).cpu()
    o = Planes(2, 6, 9, 64, True, dev)
    L.check(L.lib().rrv_maxpool2x2(L.ptr(xp.hi), L.ptr(xp.lo), 2, 13, 18, 64, L.ptr(o.hi), L.ptr(o.lo), L.stream()))
    assert torch.equal(_from_planes(L, o).cpu(), F.max_pool2d(xq, 2, 2))


@pytest.mark.parametrize("shape", [(3, 10, 12, 64), (2, 7, 9, 512), (1, 64, 64, 128), (5, 3, 3, 256)])
def test_saved_stat_tables_match_instance_norm_compute(L, dev, shape):
    from oracle import stylenet
    g = torch.Generator().manual_seed(3)
    x = torch.randn(*shape, generator=g) * 3 + 1.5           # NHWC
    st, _ = stylenet.in_compute(x.permute(0, 3, 1, 2))
    ref = torch.stack([t.reshape(-1) for t in st])
    Cc = shape[-1]
    xd = x.to(dev).contiguous()
    part = torch.empty((5, Cc), dtype=torch.float64, device=dev)
    out = torch.empty((4, Cc), dtype=torch.float32, device=dev)
    L.check(L.lib().rrv_channel_stats(xd.data_ptr(), xd.numel() // Cc, Cc, part.data_ptr(), L.stream()))
    L.check(L.lib().rrv_stats_finalize(part.data_ptr(), Cc, 0, 1e-8, out.data_ptr(), L.stream()))
    got = out.cpu()
    for r in range(4):
        assert rel_linf(got[r], ref[r]) < 1e-5, r
    # unbiased mean/std of EncoderStyle.cal_mean_std on the first sample
    ms = stylenet.cal_mean_std(x[:1].permute(0, 3, 1, 2))
    x1 = xd[:1].contiguous()
    L.check(L.lib().rrv_channel_stats(x1.data_ptr(), x1.numel() // Cc, Cc, part.data_ptr(), L.stream()))
    out2 = torch.empty((2, Cc), dtype=torch.float32, device=dev)
    L.check(L.lib().rrv_stats_finalize(part.data_ptr(), Cc, 1, 1e-5, out2.data_ptr(), L.stream()))
    assert rel_linf(out2[0].cpu(), ms.std.reshape(-1)) < 1e-5 and rel_linf(out2[1].cpu(), ms.mean.reshape(-1)) < 1e-5


def test_stats_merge_equals_single_pass(L, dev):
    """Sharded pre-pass: per-rank partials merged with Chan's formula == statistics of the whole batch."""
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(6, 8, 8, 128, generator=g) * 2 + 0.3).to(dev)
    Cc = 128
    whole = torch.empty((5, Cc), dtype=torch.float64, device=dev)
    L.check(L.lib().rrv_channel_stats(x.data_ptr(), x.numel() // Cc, Cc, whole.data_ptr(), L.stream()))
    parts = torch.empty((3, 5, Cc), dtype=torch.float64, device=dev)
    for i, (a, b) in enumerate(((0, 1), (1, 4), (4, 6))):
        xs = x[a:b].contiguous()
        L.check(L.lib().rrv_channel_stats(xs.data_ptr(), xs.numel() // Cc, Cc, parts[i].data_ptr(), L.stream()))
    merged = torch.empty_like(whole)
    L.check(L.lib().rrv_stats_merge(parts.data_ptr(), 3, Cc, merged.data_ptr(), L.stream()))
    assert torch.equal(merged[0], whole[0]) and torch.equal(merged[3:], whole[3:])
    # sums: double atomics in a different order; M2: both sides square fp32 differences about (different) fp32 means
    assert torch.allclose(merged[1], whole[1], rtol=1e-10, atol=1e-9) and torch.allclose(merged[2], whole[2], rtol=1e-5)


# ---------------------------------------------------------------------------------- whole path

def _run_cuda_global(name, state_dict, dev, precision="x3", impl="auto"):
    from oracle import cases
    from rerevst_code_b200.style_network_global import TransformerNet
    style, samples, frame = cases.global_inputs(name)
    net = TransformerNet(precision=precision, impl=impl).to(dev)
    net.load_state_dict(state_dict)
    net.generate_style_features(style.to(dev))
    net.clean()
    for s in samples:
        net.add(s.to(dev))
    net.compute()
    out = net(frame.to(dev))
    return net, out


@pytest.mark.parametrize("impl", ["ffma", "tc"])
@pytest.mark.parametrize("name", ["small_q3", "n1", "cfg1_256"])
def test_global_mode_matches_reference_golden(L, dev, state_dict, name, impl):
    if impl == "tc" and len(_impls(L)) < 2:
        pytest.skip("tcgen05 path not built")
    gold = np.load(os.path.join(GOLDEN, f"global_{name}.npz"))
    net, out = _run_cuda_global(name, state_dict, dev, impl=impl)
    eng = net._engine
    for lvl in ("relu1_1", "relu2_1", "relu3_1", "relu4_1"):
        tab = eng.style["tabs"][lvl].cpu().numpy()            # {std, mean}
        assert rel_linf(tab[1], gold[f"style/{lvl}"][0]) < TOL and rel_linf(tab[0], gold[f"style/{lvl}"][1]) < TOL
    for k in eng.stats:
        got = eng.stats[k].cpu().numpy()
        for r in range(4):
            assert rel_linf(got[r], gold["stat/" + k][r]) < TOL, (k, r)
    for f in ("Filter1", "Filter2", "Filter3"):
        for j, p in enumerate(("F1", "F2")):
            assert rel_linf(eng.filters[f][j].cpu().numpy(), gold[f"filter/{f}.{p}"]) < TOL, (f, p)
    assert tuple(out.shape) == gold["out"].shape
    assert rel_linf(out.cpu().numpy(), gold["out"]) < TOL


def test_frame_mode_matches_reference_golden(L, dev, state_dict):
    """use_Global=False (test/style_network_frame.py): per-frame statistics and per-frame dynamic filters."""
    from oracle import cases
    from rerevst_code_b200.style_network_frame import TransformerNet
    gold = np.load(os.path.join(GOLDEN, "frame_frame_small.npz"))
    style, frame = cases.frame_inputs("frame_small")
    net = TransformerNet().to(dev)
    net.load_state_dict(state_dict)
    net.generate_style_features(style.to(dev))
    out = net(frame.to(dev)).cpu().numpy()
    assert rel_linf(out, gold["out"]) < TOL
    two = net(torch.cat([frame, frame.flip(3)], 0).to(dev)).cpu().numpy()          # a batch = independent frames
    assert rel_linf(two[0], gold["out"][0]) < TOL and not np.allclose(two[1], two[0])
    with pytest.raises(AttributeError):
        net.add(frame.to(dev))


def test_train_model_validation_and_vgg_losses(L, dev, state_dict):
    """train/style_networks.py on the temporal-loss path (train.py:375-388): validation(), Vgg19 features, calc_mean_std,
    style_loss, content_loss against the unmodified reference's outputs (tests/golden/train_model.npz)."""
    from oracle import cases
    from rerevst_code_b200.style_networks import TransformerNet
    gold = np.load(os.path.join(GOLDEN, "train_model.npz"))
    style, frame = cases.frame_inputs("frame_small")
    net = TransformerNet().to(dev)
    net.load_state_dict(state_dict)
    out = net.validation(frame.to(dev), style.to(dev)).cpu().numpy()
    assert rel_linf(out, gold["validation"]) < TOL
    g = torch.Generator().manual_seed(4321)
    other = torch.randn(2, 3, 40, 56, generator=g)
    fa = net.vgg19(other.to(dev))
    fb = net.vgg19(torch.flip(other, dims=(0, 3)).to(dev))
    assert rel_linf(fa.relu4_1.cpu().numpy(), gold["relu4_1"]) < TOL
    for lvl, ft in zip(fa._fields, fa):
        mean, std = net.calc_mean_std(ft)
        assert rel_linf(mean.cpu().numpy(), gold[f"mean/{lvl}"]) < TOL and rel_linf(std.cpu().numpy(), gold[f"std/{lvl}"]) < TOL
    sl, cl = float(net.style_loss(fa, fb)), float(net.content_loss(fa, fb))
    assert abs(sl - float(gold["style_loss"])) < 2e-3 * float(gold["style_loss"])
    assert abs(cl - float(gold["content_loss"])) < 2e-3 * float(gold["content_loss"])


def test_stylization_frame_mode_facade(L, dev, state_dict):
    """framework.Stylization(use_Global=False): transfer() works without add / compute, which raise like the reference."""
    from oracle import stylenet
    from rerevst_code_b200.framework import Stylization
    rng = np.random.RandomState(5)
    style = rng.randint(0, 256, (48, 56, 3)).astype(np.uint8)
    frame = rng.randint(0, 256, (40, 64, 3)).astype(np.uint8)
    fw = Stylization(state_dict, cuda=True, use_Global=False)
    fw.prepare_style(style)
    got = fw.transfer(frame)
    fs = stylenet.encoder_style(stylenet.transform_image(stylenet.numpy2tensor(style)), state_dict)
    ref = stylenet.frame_mode_forward(state_dict, stylenet.transform_image(stylenet.numpy2tensor(frame)), fs)
    ref = stylenet.tensor2numpy(stylenet.transform_back_image(ref))
    assert got.shape == ref.shape and np.abs(got - ref).max() < 255 * TOL * 4
    with pytest.raises(AttributeError):
        fw.add(frame)


def test_full_size_raw_1080p_against_oracle(L, dev, state_dict):
    """BASELINE's full frame size WITHOUT the script's padding: 1080x1920 is 135x240 at 1/8 scale -- no multiple of any tile
    (ragged row tiles at every level, the fused max-pools next to image edges).  The CPU oracle takes the clip statistics
    from the GPU pre-pass, so the comparison isolates the per-frame forward; ~4 s of CPU time."""
    import bench
    from oracle import stylenet
    from rerevst_code_b200.framework import Stylization
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(bench.synthetic_frame(256, 320, 1))
    fw.clean()
    for i in range(2):
        fw.add(bench.synthetic_frame(270, 480, 50 + i))
    fw.compute()
    eng = fw.model._eng()
    frame = bench.synthetic_frame(1080, 1920, 100)
    got = eng.forward(torch.from_numpy(frame).unsqueeze(0).to(dev), kind=1).cpu()
    o = stylenet.GlobalOracle(state_dict)
    st = eng.export_clip_state()
    clip = stylenet.ClipState()
    for k, v in st["stats"].items():
        t = v.cpu()
        clip.stats[k] = stylenet.SavedStat(*[t[i].view(1, -1, 1, 1) for i in range(4)])
    for k, (a, b) in st["filters"].items():
        clip.filters[k] = (a.cpu().view(1, 32, 32), b.cpu().view(1, 32, 32))
    o.clip = clip
    ms_ = {k: stylenet.MeanStd(v[1].cpu().view(1, -1, 1, 1), v[0].cpu().view(1, -1, 1, 1)) for k, v in eng.style["tabs"].items()}
    o.F_style = stylenet.StyleFeatures(None, ms_["relu1_1"], ms_["relu2_1"], ms_["relu3_1"], ms_["relu4_1"])
    torch.set_num_threads(os.cpu_count() or 1)
    ref = o.forward(stylenet.transform_image(stylenet.numpy2tensor(frame)))
    assert tuple(got.shape) == (1, 3, 1080, 1920)
    assert rel_linf(got.numpy(), ref.numpy()) < TOL


def test_multi_style_interpolation_matches_reference_golden(L, dev, state_dict):
    """multi_style.TransformerNet against the unmodified "Multi-style Interpolation/style_network.py" (tests/golden/multi_style.npz)."""
    from oracle.make_golden import multi_inputs
    from rerevst_code_b200.multi_style import TransformerNet
    gold = np.load(os.path.join(GOLDEN, "multi_style.npz"))
    styles, patches, frame = multi_inputs()
    net = TransformerNet(style_num=2).to(dev)
    net.load_state_dict(state_dict)
    for i, s in enumerate(styles):
        net.generate_style_features(s.to(dev), i)
    net.clean()
    for pch in patches:
        net.add_patch(net.generate_content_features(pch.to(dev)))
    net.compute_norm()
    fc = net.generate_content_features(frame.to(dev))
    for name, w in (("w10", [1.0, 0.0]), ("w37", [0.3, 0.7]), ("w55", [0.5, 0.5])):
        assert rel_linf(net(fc, w).cpu().numpy(), gold["out/" + name]) < TOL, name
    with pytest.raises(ValueError):
        net(fc, [1.0])


def test_forward_with_oracle_statistics(L, dev, state_dict):
    """Per-frame forward alone: clip state imported from the oracle, so only forward error counts."""
    from oracle import cases, stylenet
    from rerevst_code_b200.style_network_global import TransformerNet
    style, samples, frame = cases.global_inputs("small_q3")
    o = stylenet.GlobalOracle(state_dict)
    o.generate_style_features(style)
    o.clean()
    for s in samples:
        o.add(s)
    o.compute()
    ref = o.forward(frame)
    net = TransformerNet().to(dev)
    net.load_state_dict(state_dict)
    net.generate_style_features(style.to(dev))
    eng = net._engine
    stats = {k: torch.stack([t.reshape(-1) for t in v]).contiguous() for k, v in o.clip.stats.items()}
    filters = {k: (a.reshape(32, 32).contiguous(), b.reshape(32, 32).contiguous()) for k, (a, b) in o.clip.filters.items()}
    eng.import_clip_state(dict(stats=stats, filters=filters))
    for impl in [n for n, _ in _impls(L)]:
        eng.impl_name = impl
        assert rel_linf(net(frame.to(dev)).cpu(), ref) < TOL, impl
    # Q2: a batch of frames == independent B=1 forwards
    eng.impl_name = "auto"
    g = torch.Generator().manual_seed(8)
    f2 = torch.randn(1, 3, 64, 96, generator=g)
    both = net(torch.cat([frame, f2]).to(dev)).cpu()
    assert rel_linf(both[:1], ref) < TOL and rel_linf(both[1:], o.forward(f2)) < TOL


def test_forward_before_compute_raises(dev, state_dict):
    from rerevst_code_b200.style_network_global import TransformerNet
    net = TransformerNet().to(dev)
    net.load_state_dict(state_dict)
    net.generate_style_features(torch.zeros(1, 3, 32, 32, device=dev))
    net.clean()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 32, 32, device=dev))


def test_stylization_facade_u8_path(L, dev, state_dict):
    """framework.Stylization: uint8 BGR in, float32 BGR [0,255] out, incl. the script's crop."""
    from oracle import stylenet
    from rerevst_code_b200.framework import Stylization
    rng = np.random.RandomState(0)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    style, f0, f1, frame = smooth(64, 72), smooth(48, 64), smooth(48, 64), smooth(64, 96)
    o = stylenet.GlobalOracle(state_dict)
    o.generate_style_features(stylenet.transform_image(stylenet.numpy2tensor(style)))
    o.clean()
    for f in (f0, f1):
        o.add(stylenet.transform_image(stylenet.numpy2tensor(f)))
    o.compute()
    ref = o.transfer(frame)
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(style)
    fw.clean()
    fw.add(f0)
    fw.add(f1)
    fw.compute()
    got = fw.transfer(frame)
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert float(np.max(np.abs(got - ref))) < 255 * TOL
    crop = fw.transfer(frame, crop=(8, 16, 40, 64))
    assert np.array_equal(crop, got[8:48, 16:80])


# ---------------------------------------------------------------------------------- warp

def _index_image(b, h, w):
    img = np.zeros((b, 2, h, w), np.float32)
    img[:, 0] = np.arange(w, dtype=np.float32)[None, None, :]
    img[:, 1] = np.arange(h, dtype=np.float32)[None, :, None]
    return img


def test_warp_indices_bit_exact(dev):
    from oracle import cases, warp as ow
    from rerevst_code_b200.loss_networks import warp, warp_indices
    with open(os.path.join(GOLDEN, "warp_digests.json")) as f:
        digests = json.load(f)
    for (h, w) in cases.WARP_SIZES + cases.WARP_SMALL:
        flo = cases.warp_flow(h, w)
        idx = warp_indices(torch.from_numpy(flo).to(dev)).cpu().numpy()
        iy, ix = ow.warp_indices(flo)
        assert np.array_equal(idx[..., 0], iy) and np.array_equal(idx[..., 1], ix), (h, w)
        assert cases.digest(idx[..., 1].astype(np.int32)) == digests[f"{h}x{w}"]["ix_sha256"]
        assert cases.digest(idx[..., 0].astype(np.int32)) == digests[f"{h}x{w}"]["iy_sha256"]
        img = _index_image(flo.shape[0], h, w)
        out = warp(torch.from_numpy(img).to(dev), torch.from_numpy(flo).to(dev)).cpu().numpy()
        assert np.array_equal(out, ow.warp(img, flo))


def test_temporal_loss_and_backward(dev):
    from oracle import cases, warp as ow
    from rerevst_code_b200.loss_networks import TemporalLoss, warp
    gold = np.load(os.path.join(GOLDEN, "warp.npz"))
    g = torch.Generator().manual_seed(99)
    first = torch.randn(2, 3, 64, 64, generator=g)
    second = torch.randn(2, 3, 64, 64, generator=g)
    flo = torch.from_numpy(cases.warp_flow(64, 64))
    loss, warped = TemporalLoss()(first.to(dev), second.to(dev), flo.to(dev))
    assert np.array_equal(warped.cpu().numpy(), gold["tl/warped"])
    assert abs(float(loss) - float(gold["tl/loss"])) <= 1e-6 * float(gold["tl/loss"])
    # backward: scatter-add of the output gradient to the source pixels
    x = first.to(dev).requires_grad_(True)
    go = torch.randn(2, 3, 64, 64, generator=g)
    warp(x, flo.to(dev)).backward(go.to(dev))
    iy, ix = ow.warp_indices(flo.numpy())
    ref = np.zeros((2, 3, 64, 64), np.float64)
    for b in range(2):
        for c in range(3):
            np.add.at(ref[b, c], (iy[b], ix[b]), go[b, c].numpy().astype(np.float64))
    assert rel_linf(x.grad.cpu().numpy(), ref) < 1e-5


# ---------------------------------------------------------------------------------- kernel variants / entry point

@pytest.mark.parametrize("tune", [(256, 1, 1, 1), (128, 2, 1, 1), (256, 2, 0, 1), (256, 2, 1, 0), (64, 1, 0, 0)])
def test_conv_main_loop_variants(L, dev, tune):
    """rrv_tc_tune / rrv_tc_tune_pair / rrv_tc_tune_merge: one M tile per weight tile, narrower Cout tiles, no CTA pairs, no merged
    taps (the defaults -- 256-wide tiles, two M tiles, pairs, merged taps -- run in every other test)."""
    from rerevst_code_b200.engine import ConvW, make_epilogue
    L.check(L.lib().rrv_tc_tune(tune[0], tune[1]))
    L.check(L.lib().rrv_tc_tune_pair(tune[2], 128))
    L.check(L.lib().rrv_tc_tune_merge(tune[3]))
    try:
        for case in [(1, 40, 24, 64, 64, 3, 0), (2, 36, 20, 128, 128, 3, 0), (1, 48, 32, 128, 64, 3, 1), (1, 40, 16, 256, 256, 3, 1),
                     (1, 33, 17, 64, 128, 1, 0)]:
            N, H, W, Cin, Cout, k, ups = case
            g = torch.Generator().manual_seed(7)
            hin, win = (H // 2, W // 2) if ups else (H, W)
            x = torch.randn(N, Cin, hin, win, generator=g)
            w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
            b = torch.randn(Cout, generator=g) * 0.1
            cw = ConvW(w.to(dev), b.to(dev), ups=bool(ups))
            xp = _to_planes(L, x.to(dev))
            xq = _from_planes(L, xp).cpu()
            xin = F.interpolate(xq, scale_factor=2, mode="nearest") if ups else xq
            ref = F.leaky_relu(F.conv2d(xin, w, b, padding=k // 2), 0.2)
            d = L.Conv()
            d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, k, ups
            d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
            d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
            d.ep = make_epilogue(bias=cw.bias, act=2)
            d.out_mode = L.OUT_F32_NHWC
            out = torch.full((N, H, W, Cout), float("nan"), dtype=torch.float32, device=dev)
            d.out_f32 = out.data_ptr()
            L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()), str(case))
            assert rel_linf(out.permute(0, 3, 1, 2).cpu(), ref) < 2e-4, (tune, case)
    finally:
        L.check(L.lib().rrv_tc_tune(256, 2))
        L.check(L.lib().rrv_tc_tune_pair(1, 64))
        L.check(L.lib().rrv_tc_tune_merge(1))


@pytest.mark.parametrize("case", [(1, 64, 40, 64, 64), (2, 37, 21, 64, 64), (1, 32, 48, 128, 128), (2, 19, 27, 128, 128),
                                  (1, 24, 16, 256, 256), (1, 60, 8, 64, 96)])
@pytest.mark.parametrize("pair,merge", [(1, 1), (0, 1), (1, 0)])
def test_conv_fused_maxpool(L, dev, case, pair, merge):
    """rrv_conv.pool: nn.MaxPool2d(2, 2) of vgg19.features[4|9|18] in the epilogue of the convolution before it
    (merged-tap layout: column partner exchanged between warps; plain layout: both partners in the warp).
    Must equal the separate pool kernel applied to the unpooled kernel output bit for bit (odd sizes floor)."""
    from rerevst_code_b200.engine import ConvW, Planes, make_epilogue
    N, H, W, Cin, Cout = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    cw = ConvW(w.to(dev), b.to(dev))
    xp = _to_planes(L, x.to(dev))
    L.check(L.lib().rrv_tc_tune_pair(pair, 64))
    L.check(L.lib().rrv_tc_tune_merge(merge))
    try:
        outs = []
        for pool in (0, 1):
            d = L.Conv()
            d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, 3, 0
            d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
            d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
            d.ep = make_epilogue(bias=cw.bias, act=1)
            d.pool = pool
            o = Planes(N, H >> pool, W >> pool, Cout, True, dev)
            o.hi.fill_(float("nan"))
            d.out_mode, d.out_hi, d.out_lo = L.OUT_PLANES, L.ptr(o.hi), L.ptr(o.lo)
            L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()), str(case))
            outs.append(_from_planes(L, o).cpu())
        ref = F.relu(F.conv2d(_from_planes(L, xp).cpu(), w, b, padding=1))
        assert rel_linf(outs[0], ref) < 2e-4
        assert torch.equal(outs[1], F.max_pool2d(outs[0], 2, 2)), (case, pair, merge)
    finally:
        L.check(L.lib().rrv_tc_tune_pair(1, 64))
        L.check(L.lib().rrv_tc_tune_merge(1))


def test_reflect_pad_matches_copy_make_border(L, dev):
    """rrv_reflect_pad_u8 = cv2.copyMakeBorder(..., BORDER_REFLECT) of ReshapeTool.process (edge pixel repeated), including
    borders wider than the image; transfer_stream(pad_to=...) equals padding on the host."""
    rng = np.random.RandomState(9)
    for (h, w, ph, pw) in ((40, 56, 192, 192), (3, 5, 192, 256), (1080, 1920, 1216, 2048)):
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = np.pad(img, ((64, ph - 64 - h), (64, pw - 64 - w), (0, 0)), mode="symmetric")
        try:
            import cv2
            assert np.array_equal(ref, cv2.copyMakeBorder(img, 64, ph - 64 - h, 64, pw - 64 - w, cv2.BORDER_REFLECT))
        except ImportError:
            pass
        src = torch.from_numpy(img).unsqueeze(0).to(dev)
        dst = torch.empty((1, ph, pw, 3), dtype=torch.uint8, device=dev)
        L.check(L.lib().rrv_reflect_pad_u8(src.data_ptr(), 1, h, w, 64, 64, ph, pw, dst.data_ptr(), L.stream()))
        assert np.array_equal(dst[0].cpu().numpy(), ref), (h, w)


def test_transfer_stream_equals_transfer(L, dev, state_dict):
    from rerevst_code_b200.framework import Stylization
    rng = np.random.RandomState(5)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(smooth(64, 64))
    fw.clean()
    fw.add(smooth(48, 64))
    fw.compute()
    frames = [smooth(64, 128) for _ in range(7)]
    crop = (8, 16, 40, 96)
    want = [fw.transfer(f, crop=crop) for f in frames]
    got = list(fw.transfer_stream(iter(frames), crop=crop, depth=3))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    for i, a in enumerate(fw.transfer_stream(iter(frames), crop=crop, depth=2, copy=False)):     # zero-copy: valid until the next item
        assert np.array_equal(a, want[i])
    for lanes in (1, 2, 3):              # frames in flight on separate compute streams: same frames, same order
        for out_dtype in ("f32", "u8"):
            got = list(fw.transfer_stream(iter(frames), crop=crop, lanes=lanes, out_dtype=out_dtype))
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert np.array_equal(a, b if out_dtype == "f32" else np.rint(b).astype(np.uint8)), (lanes, out_dtype)
    # pad_to: raw frames, reflect border on the device, cropped back to the raw window (generate_real_video.py:66-83, :167)
    raw = [smooth(40, 56) for _ in range(4)]
    padded = [np.pad(f, ((64, 192 - 64 - 40), (64, 192 - 64 - 56), (0, 0)), mode="symmetric") for f in raw]
    want = [fw.transfer(f, crop=(64, 64, 40, 56)) for f in padded]
    got = list(fw.transfer_stream(iter(raw), pad_to=(192, 192)))
    for a, b in zip(got, want):
        assert a.shape == (40, 56, 3) and np.array_equal(a, b)


def test_forward_graphed_lanes_run_concurrently_and_agree(L, dev, state_dict):
    """Two frames in flight (engine.forward_graphed(..., lane=i) on engine.lane_streams()): every lane returns bit for bit what
    a single lane returns, whatever the interleaving on the device."""
    from rerevst_code_b200.framework import Stylization
    g = torch.Generator().manual_seed(3)
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(torch.randint(0, 256, (64, 64, 3), generator=g, dtype=torch.uint8).numpy())
    fw.clean()
    fw.add(torch.randint(0, 256, (48, 64, 3), generator=g, dtype=torch.uint8).numpy())
    fw.compute()
    eng = fw.model._eng()
    frames = [torch.randint(0, 256, (1, 192, 256, 3), generator=g, dtype=torch.uint8).to(dev) for _ in range(6)]
    post = ("u8", (8, 8, 160, 200))
    want = [eng.forward_graphed(f, kind=1, post=post).clone() for f in frames]
    streams = eng.lane_streams(2)
    outs = [None] * len(frames)
    for rep in range(3):
        for i, f in enumerate(frames):
            with torch.cuda.stream(streams[i % 2]):
                outs[i] = eng.forward_graphed(f, kind=1, post=post, lane=i % 2).clone()
        for st in streams[1:]:
            streams[0].wait_stream(st)
        torch.cuda.synchronize()
        for a, b in zip(outs, want):
            assert torch.equal(a, b)


def test_generate_real_video_entry(tmp_path, dev, state_dict):
    """The packaged generate_real_video.main against the same steps done with the CPU oracle."""
    import cv2
    from oracle import stylenet
    from rerevst_code_b200 import generate_real_video as grv
    rng = np.random.RandomState(9)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    vid = tmp_path / "inputs" / "clip"
    vid.mkdir(parents=True)
    frames = []
    for i in range(10):
        f = smooth(40, 56)
        frames.append(f)
        cv2.imwrite(str(vid / f"frame_{i:04d}.png"), f)
    style = smooth(64, 72)
    cv2.imwrite(str(tmp_path / "style.png"), style)
    ckpt = tmp_path / "net.pth"
    torch.save(state_dict, str(ckpt))
    out_dir = grv.main(style_img=str(tmp_path / "style.png"), content_video=str(vid / "*.png"), checkpoint_path=str(ckpt),
                       result_frames_path=str(tmp_path / "rf"), result_videos_path=str(tmp_path / "rv"), save_video=False,
                       verbose=False)
    import glob
    order = glob.glob(str(vid / "*.png"))                        # the script's (unsorted) glob order
    o = stylenet.GlobalOracle(state_dict)
    o.generate_style_features(stylenet.transform_image(stylenet.numpy2tensor(cv2.imread(str(tmp_path / "style.png")))))
    o.clean()
    for s in range((10 - 1) // 8):
        o.add(stylenet.transform_image(stylenet.numpy2tensor(cv2.imread(order[s * 8]))))
    o.add(stylenet.transform_image(stylenet.numpy2tensor(cv2.imread(order[-1]))))
    o.compute()
    tool = grv.ReshapeTool()
    for path in order:
        img = cv2.imread(path)
        ref = o.transfer(tool.process(img))[64:64 + 40, 64:64 + 56]
        got = cv2.imread(os.path.join(out_dir, os.path.basename(path))).astype(np.float32)
        assert got.shape == ref.shape
        assert float(np.max(np.abs(got - np.clip(np.rint(ref), 0, 255)))) <= 1.0


# ---------------------------------------------------------------------------------- round 2: sizes / paths the bench times

def _oracle_from_engine(state_dict, eng):
    import bench
    return bench.oracle_from_engine(state_dict, eng)


def test_frame_mode_ragged_270x480_against_oracle(L, dev, state_dict):
    """Frame mode (use_Global=False) on a frame that is no multiple of 8 rows (270 -> 33.75 at 1/8 scale): per-frame
    statistics, per-frame filters, output 264 x 480 like the reference (three floor pools, three x2 upsamples)."""
    import bench
    from oracle import stylenet
    from rerevst_code_b200.framework import Stylization
    style, frame = bench.synthetic_frame(96, 128, 3), bench.synthetic_frame(270, 480, 4)
    fw = Stylization(state_dict, cuda=True, use_Global=False)
    fw.prepare_style(style)
    got = fw.transfer(frame)
    fs = stylenet.encoder_style(stylenet.transform_image(stylenet.numpy2tensor(style)), state_dict)
    ref = stylenet.frame_mode_forward(state_dict, stylenet.transform_image(stylenet.numpy2tensor(frame)), fs)
    ref = stylenet.tensor2numpy(stylenet.transform_back_image(ref))
    assert got.shape == ref.shape == (264, 480, 3)
    assert float(np.abs(got - ref).max()) < 255 * TOL
    u8 = fw.transfer(frame, out_dtype="u8")
    assert u8.dtype == np.uint8 and np.array_equal(u8, np.rint(got).astype(np.uint8))


def test_prepass_8_samples_540x960_against_oracle(L, dev, state_dict):
    """Decoder.compute (style_network_global.py:425-439) at a realistic size: 8 sampled frames of 540 x 960 (ragged at 1/8
    scale: 67.5 x 120), all 11 statistic tables and 6 dynamic filters against the CPU oracle's own pre-pass (~20 s of CPU)."""
    import bench
    from oracle import stylenet
    from rerevst_code_b200.framework import Stylization
    torch.set_num_threads(os.cpu_count() or 1)
    style = bench.synthetic_frame(256, 320, 1)
    frames = [bench.synthetic_frame(540, 960, 60 + i) for i in range(8)]
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(style)
    fw.clean()
    for f in frames:
        fw.add(f)
    fw.compute()
    eng = fw.model._eng()
    o = stylenet.GlobalOracle(state_dict)
    o.generate_style_features(stylenet.transform_image(stylenet.numpy2tensor(style)))
    o.clean()
    for f in frames:
        o.add(stylenet.transform_image(stylenet.numpy2tensor(f)))
    o.compute()
    for k, st in o.clip.stats.items():
        got = eng.stats[k].cpu().numpy()
        for r in range(4):
            assert rel_linf(got[r], st[r].reshape(-1).numpy()) < TOL, (k, r)
    for f, (a, b) in o.clip.filters.items():
        assert rel_linf(eng.filters[f][0].cpu().numpy(), a.reshape(32, 32).numpy()) < TOL, f
        assert rel_linf(eng.filters[f][1].cpu().numpy(), b.reshape(32, 32).numpy()) < TOL, f
    # and a frame through both with their OWN pre-pass results: the whole script path at this size
    frame = bench.synthetic_frame(540, 960, 99)
    assert float(np.abs(fw.transfer(frame) - o.transfer(frame)).max()) < 255 * TOL


def test_non_multiple_of_8_frame_436x1024(L, dev, state_dict):
    """A raw Sintel-sized frame (test/inputs/ambush_4: 436 x 1024): the reference returns 432 x 1024; every entry point
    (eager forward, graph replay, transfer, transfer_stream, fused u8 head) must agree with the oracle on that shape."""
    import bench
    from rerevst_code_b200.framework import Stylization
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(bench.synthetic_frame(128, 160, 1))
    fw.clean()
    for i in range(2):
        fw.add(bench.synthetic_frame(218, 512, 70 + i))
    fw.compute()
    eng = fw.model._eng()
    frame = bench.synthetic_frame(436, 1024, 71)
    o = _oracle_from_engine(state_dict, eng)
    ref = o.transfer(frame)
    assert ref.shape == (432, 1024, 3)
    d = torch.from_numpy(frame).unsqueeze(0).to(dev)
    y = eng.forward(d, kind=1)
    assert tuple(y.shape) == (1, 3, 432, 1024)
    yg = eng.forward_graphed(d, kind=1)
    assert tuple(yg.shape) == (1, 3, 432, 1024) and torch.equal(yg, y)
    got = fw.transfer(frame)
    assert got.shape == ref.shape and float(np.abs(got - ref).max()) < 255 * TOL
    st = list(fw.transfer_stream(iter([frame, frame[:, ::-1].copy(), frame])))
    assert all(s.shape == ref.shape for s in st) and np.array_equal(st[0], got) and np.array_equal(st[2], got)
    with pytest.raises(ValueError):
        fw.transfer(frame, crop=(0, 0, 436, 1024))                     # the window must lie inside the 432 x 1024 result
    with pytest.raises(ValueError):
        eng.forward(d, kind=1, out=torch.empty((1, 3, 436, 1024), device=dev))


def test_head_fused_postprocess_and_u8(L, dev, state_dict):
    """The RGB head writing the finished frame (RRV_OUT_BGR_F32 / _U8: transform_back_image + tensor2numpy + crop in the epilogue)
    against the separate rrv_postprocess_bgr kernel on the head's NCHW output: bit-equal; uint8 = rint of the float frame."""
    import bench
    from rerevst_code_b200.framework import Stylization
    fw = Stylization(state_dict, cuda=True)
    fw.prepare_style(bench.synthetic_frame(96, 96, 1))
    fw.clean()
    fw.add(bench.synthetic_frame(64, 96, 2))
    fw.compute()
    eng = fw.model._eng()
    d = torch.from_numpy(bench.synthetic_frame(104, 200, 3)).unsqueeze(0).to(dev)
    y = eng.forward(d, kind=1)
    for crop in (None, (8, 24, 80, 150), (0, 0, 1, 1), (103, 199, 1, 1)):
        want = eng.postprocess(y, crop, "f32")
        got = eng.forward(d, kind=1, post=("f32", crop))
        assert got.dtype == torch.float32 and torch.equal(got, want), crop
        u8 = eng.forward(d, kind=1, post=("u8", crop))
        assert u8.dtype == torch.uint8 and torch.equal(u8, torch.round(want).to(torch.uint8)), crop
        assert torch.equal(eng.postprocess(y, crop, "u8"), u8)
        assert torch.equal(eng.forward_graphed(d, kind=1, post=("u8", crop)), u8)
    assert float(want.min()) >= 0.0 and float(want.max()) <= 255.0


def test_conv_operand_terms_knob(L, dev):
    """rrv_conv.terms: NO_WLO multiplies full activations by bf16-rounded weights, NO_ALO bf16-rounded activations by full
    weights (two MMAs per k-slice instead of three); each must match exactly that arithmetic on the CPU."""
    from rerevst_code_b200.engine import ConvW, make_epilogue
    if len(_impls(L)) < 2:
        pytest.skip("tcgen05 path not built")
    g = torch.Generator().manual_seed(77)
    bf = lambda t: t.to(torch.bfloat16).float()
    two = lambda t: bf(t) + bf(t - bf(t))
    for (N, H, W, Cin, Cout, k) in ((1, 24, 70, 64, 64, 3), (1, 20, 24, 512, 64, 3), (2, 16, 16, 128, 256, 3)):
        x = torch.randn(N, Cin, H, W, generator=g)
        w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        cw = ConvW(w.to(dev), None)
        xp = _to_planes(L, x.to(dev))
        refs = {L.TERMS_FULL: F.conv2d(two(x), two(w), padding=1), L.TERMS_NO_WLO: F.conv2d(two(x), bf(w), padding=1),
                L.TERMS_NO_ALO: F.conv2d(bf(x), two(w), padding=1)}
        outs = {}
        for terms, ref in refs.items():
            d = L.Conv()
            d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, k, 0
            d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
            d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
            d.ep = make_epilogue()
            d.out_mode = L.OUT_F32_NHWC
            d.terms = terms
            out = torch.empty((N, H, W, Cout), dtype=torch.float32, device=dev)
            d.out_f32 = out.data_ptr()
            L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()))
            outs[terms] = out.permute(0, 3, 1, 2).cpu()
            assert rel_linf(outs[terms], ref) < 3e-5, (terms, Cin, Cout)
        assert rel_linf(outs[L.TERMS_NO_WLO], refs[L.TERMS_FULL]) > 1e-4          # the dropped term is really dropped


def test_temporal_loss_config5_512_b4(dev):
    """BASELINE.json configs[4]: B = 4 pairs at 512 x 512, the flow of GenerateFakeFlow (train/loss_networks.py:71-86) under
    fixed seeds; warp bit-equal to the oracle, loss within 1e-6; GenerateFakeData (:88-104) = warp + N(0, sigma in [1e-3, 2e-3])."""
    import random
    from oracle import warp as ow
    from rerevst_code_b200.loss_networks import TemporalLoss, warp, warp_indices
    B, Cc, H, W = 4, 3, 512, 512
    g = torch.Generator().manual_seed(5)
    first = torch.randn(B, Cc, H, W, generator=g)
    tl = TemporalLoss()
    np.random.seed(0)
    random.seed(0)
    flow = tl.GenerateFakeFlow(H, W)
    assert tuple(flow.shape) == (2, H, W) and float(flow.abs().max()) < 40.0
    flow_b = flow.unsqueeze(0).expand(B, 2, H, W).contiguous()
    iy, ix = ow.warp_indices(flow_b.numpy())
    idx = warp_indices(flow_b.to(dev)).cpu().numpy()
    assert np.array_equal(idx[..., 0], iy) and np.array_equal(idx[..., 1], ix)
    ref_w = ow.warp(first.numpy(), flow_b.numpy())
    second = torch.from_numpy(ref_w) + 1.5e-3 * torch.randn(B, Cc, H, W, generator=g)
    loss, warped = tl(first.to(dev), second.to(dev), flow_b.to(dev))
    assert np.array_equal(warped.cpu().numpy(), ref_w)
    ref_loss = float(np.mean(np.abs(ref_w.astype(np.float64) - second.numpy().astype(np.float64))))
    assert abs(float(loss) - ref_loss) <= 1e-6 * ref_loss
    # GenerateFakeData: same seeds -> same flow; second = warp(first, flow) + Gaussian noise of sigma in [noise_level, 2 noise_level)
    np.random.seed(0)
    random.seed(0)
    sec, fl2 = tl.GenerateFakeData(first.to(dev))
    assert tuple(fl2.shape) == (B, 2, H, W) and torch.equal(fl2[0].cpu(), flow) and torch.equal(fl2[3].cpu(), flow)
    noise = (sec - warp(first.to(dev), fl2)).cpu()
    assert 0.9e-3 < float(noise.std()) < 2.1e-3 and abs(float(noise.mean())) < 1e-5
    # a CPU / mis-shaped second frame must not reach the fused kernel: the reference's own route is taken
    l2, _ = tl(first.to(dev), second[:, :1].to(dev), flow_b.to(dev))           # broadcasts like torch
    assert torch.isfinite(l2)
    # non-contiguous upstream gradient (channels_last): the returned gradient is in the logical NCHW order
    x = first.to(dev).requires_grad_(True)
    go = torch.randn(B, Cc, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    warp(x, flow_b.to(dev)).backward(go)
    x2 = first.to(dev).requires_grad_(True)
    warp(x2, flow_b.to(dev)).backward(go.contiguous())
    assert x.grad.is_contiguous() and torch.allclose(x.grad, x2.grad, rtol=1e-5, atol=1e-5)     # float atomics: order-dependent last bits


FUSED_STAT_CASES = [
    # N, H, W, Cin, Cout, ups  -- main-loop shapes: merged taps (Cout <= 64), CTA pairs, two Cout tiles, nearest x2, ragged edges
    (1, 22, 70, 64, 64, 0), (2, 19, 27, 128, 256, 0), (1, 24, 40, 256, 512, 0), (1, 26, 38, 256, 128, 1),
    (1, 16, 48, 128, 64, 1), (3, 9, 13, 512, 64, 0), (1, 33, 65, 128, 128, 0),
]


@pytest.mark.parametrize("case", FUSED_STAT_CASES)
def test_conv_fused_statistics(L, dev, case):
    """rrv_conv.stats: {sum, sum of squares, min, max} of the written values from the convolution's epilogue, against the
    two-pass rrv_channel_stats on the tensor it wrote (frame-mode InstanceNorm, style_network_frame.py:39-43; pre-pass :59-77)."""
    from rerevst_code_b200.engine import ConvW, make_epilogue
    if len(_impls(L)) < 2:
        pytest.skip("tcgen05 path not built")
    N, H, W, Cin, Cout, ups = case
    g = torch.Generator().manual_seed(sum(case))
    hin, win = (H // 2, W // 2) if ups else (H, W)
    x = torch.randn(N, Cin, hin, win, generator=g) + 0.3
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.2
    cw = ConvW(w.to(dev), b.to(dev), ups=bool(ups))
    xp = _to_planes(L, x.to(dev))
    for minmax in (1, 0):
        d = L.Conv()
        d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, 3, ups
        d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
        d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
        d.ep = make_epilogue(bias=cw.bias, act=2)
        d.out_mode = L.OUT_F32_NHWC
        out = torch.empty((N, H, W, Cout), dtype=torch.float32, device=dev)
        d.out_f32 = out.data_ptr()
        part = torch.empty((5, Cout), dtype=torch.float64, device=dev)
        L.check(L.lib().rrv_stats_init(part.data_ptr(), Cout, float(N * H * W), L.stream()))
        d.stats, d.stats_minmax = part.data_ptr(), minmax
        L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()))
        L.check(L.lib().rrv_stats_sums_to_m2(part.data_ptr(), Cout, L.stream()))
        plain = torch.empty_like(out)                            # the same convolution without statistics: same values
        d.stats, d.out_f32 = 0, plain.data_ptr()
        L.check(L.lib().rrv_conv2d(C.byref(d), L.IMPL_TCGEN05, L.stream()))
        # same arithmetic; the nearest-x2 layer with Cout <= 64 takes the merged-phase main loop when statistics are requested
        # (partial sums meet in the epilogue instead of in TMEM: last-bit differences)
        assert torch.equal(out, plain) or (ups and Cout <= 64 and torch.allclose(out, plain, rtol=1e-4, atol=1e-5 * float(plain.abs().max())))
        ref = torch.empty((5, Cout), dtype=torch.float64, device=dev)
        L.check(L.lib().rrv_channel_stats(out.data_ptr(), N * H * W, Cout, ref.data_ptr(), L.stream()))
        part, ref = part.cpu(), ref.cpu()
        assert torch.equal(part[0], ref[0])
        assert torch.allclose(part[1], ref[1], rtol=1e-6, atol=1e-4 * float(ref[1].abs().max()) * 1e-2)
        assert torch.allclose(part[2], ref[2], rtol=2e-5)
        if minmax:
            assert torch.equal(part[3], ref[3]) and torch.equal(part[4], ref[4])
        else:
            assert torch.isinf(part[3]).all() and torch.isinf(part[4]).all()


@pytest.mark.parametrize("shape", [(2, 9, 11, 64), (1, 17, 23, 128), (3, 5, 7, 512), (1, 40, 64, 256)])
def test_pointwise_fused_statistics(L, dev, shape):
    from rerevst_code_b200.engine import make_epilogue
    N, H, W, Cc = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(N, H, W, Cc, generator=g) * 2 + 0.5).to(dev)
    tab = torch.stack([torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5, torch.full((Cc,), -1.5), torch.full((Cc,), 2.0)]).to(dev)
    ep = make_epilogue(norm1=tab)
    out = torch.empty_like(x)
    part = torch.empty((5, Cc), dtype=torch.float64, device=dev)
    L.check(L.lib().rrv_stats_init(part.data_ptr(), Cc, float(N * H * W), L.stream()))
    L.check(L.lib().rrv_pointwise_stats(x.data_ptr(), H * W * Cc, N, H, W, Cc, C.byref(ep), L.OUT_F32_NHWC, 0, 0, out.data_ptr(),
                                        part.data_ptr(), 1, L.stream()))
    L.check(L.lib().rrv_stats_sums_to_m2(part.data_ptr(), Cc, L.stream()))
    plain = torch.empty_like(x)
    L.check(L.lib().rrv_pointwise(x.data_ptr(), H * W * Cc, N, H, W, Cc, C.byref(ep), L.OUT_F32_NHWC, 0, 0, plain.data_ptr(), L.stream()))
    assert torch.equal(out, plain)
    ref = torch.empty((5, Cc), dtype=torch.float64, device=dev)
    L.check(L.lib().rrv_channel_stats(out.data_ptr(), N * H * W, Cc, ref.data_ptr(), L.stream()))
    part, ref = part.cpu(), ref.cpu()
    assert torch.equal(part[0], ref[0]) and torch.equal(part[3], ref[3]) and torch.equal(part[4], ref[4])
    assert torch.allclose(part[1], ref[1], rtol=1e-5, atol=1e-3) and torch.allclose(part[2], ref[2], rtol=1e-4)


def test_fold_filter_kernel_matches_host_fold(L, dev, state_dict):
    """rrv_fold_filter (the KernelFilter's two 32x32 matrices folded into its convolutions on the device, straight into
    tensor-core blobs) against the same fold done with torch and packed by rrv_pack_weights_tc: the two convolutions agree."""
    from rerevst_code_b200.engine import INNER_PAD, ConvW, FoldedFilter, make_epilogue
    from rerevst_code_b200.engine import StyleEngine
    if len(_impls(L)) < 2:
        pytest.skip("tcgen05 path not built")
    eng = StyleEngine(dev)
    eng.load_weights(state_dict)
    fw = eng.w["Filter2"]
    g = torch.Generator().manual_seed(12)
    wf1, wf2 = (torch.randn(32, 32, generator=g) * 0.3).to(dev), (torch.randn(32, 32, generator=g) * 0.3).to(dev)
    down, up = FoldedFilter(fw, dev).fold(wf1, wf2)
    dw = torch.matmul(wf1, fw["down_w"].reshape(32, -1)).reshape(32, 512, 3, 3)
    db = torch.mv(wf1, fw["down_b"])
    uw = torch.einsum("ojyx,ji->oiyx", fw["up_w"], wf2).contiguous()
    down_ref, up_ref = ConvW(dw, db, cout_pad=INNER_PAD), ConvW(uw, fw["up_b"], cin_pad=INNER_PAD)
    x = _to_planes(L, torch.randn(1, 512, 20, 36, generator=g).to(dev))
    t_a = eng._conv(down, x, make_epilogue(bias=down.bias, act=2))
    t_b = eng._conv(down_ref, x, make_epilogue(bias=down_ref.bias, act=2))
    ya, yb = _from_planes(L, t_a), _from_planes(L, t_b)
    assert ya.shape[1] == INNER_PAD and (INNER_PAD == 32 or float((ya[:, 32:]).abs().max()) == 0.0) and rel_linf(ya.cpu(), yb.cpu()) < 2e-5
    u_a = eng._conv(up, t_b, make_epilogue(bias=up.bias), L.OUT_F32_NHWC)
    u_b = eng._conv(up_ref, t_b, make_epilogue(bias=up_ref.bias), L.OUT_F32_NHWC)
    assert rel_linf(u_a.cpu(), u_b.cpu()) < 2e-5


def _vgg_backward_reference(saved, cots, sd):
    """The chain rule through Vgg19 in float64 on the CPU, using the ReLU outputs the GPU forward saved (so that units whose
    pre-activation is within rounding of zero take the same side in both computations)."""
    from oracle import stylenet
    names = stylenet._enc_names("Vgg19")
    taps, pooled_after, g = {0: 0, 2: 1, 4: 2, 8: 3}, {1, 3, 7}, None
    for i in range(8, -1, -1):
        y = saved[i].permute(0, 3, 1, 2).double().cpu()
        if i in pooled_after and g is not None:
            _, idx = F.max_pool2d(y, 2, 2, return_indices=True)
            g = F.max_unpool2d(g, idx, 2, 2, output_size=y.shape[-2:])
        if i in taps and cots[taps[i]] is not None:
            g = cots[taps[i]].double() if g is None else g + cots[taps[i]].double()
        if g is None:
            continue
        g = g * (y > 0)
        g = F.conv_transpose2d(g, sd[names[i][0]].double(), padding=1)
    return g


def test_vgg19_loss_network_backward(L, dev, state_dict):
    """train.py:376-414: Loss.backward() runs first through the frozen Vgg19 (train/style_networks.py:284-314).  Its data gradient
    here = ReLU / max-pool backward kernels + the tensor-core convolution on transposed, rotated weights; checked against
    torch.autograd through the CPU oracle's Vgg19, for a weighted sum of the four features and for content + style loss."""
    from oracle import stylenet
    from rerevst_code_b200.style_networks import TransformerNet
    net = TransformerNet().to(dev)
    net.load_state_dict(state_dict)
    sd = {k: v.float() for k, v in state_dict.items()}
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 3, 44, 60, generator=g)                     # 44 -> 22 -> 11 -> 5: an odd size before the last pool
    cot = [torch.randn(s, generator=g) for s in ((2, 64, 44, 60), (2, 128, 22, 30), (2, 256, 11, 15), (2, 512, 5, 7))]
    xr = x.clone().requires_grad_(True)
    ref_feats = stylenet.vgg19_features(xr, sd, top="Vgg19")
    sum((f * c).sum() for f, c in zip(ref_feats, cot)).backward()
    xd = x.to(dev).requires_grad_(True)
    feats = net.vgg19(xd)
    for f, r in zip(feats, ref_feats):
        assert rel_linf(f.detach().cpu().numpy(), r.detach().numpy()) < TOL
    sum((f * c.to(dev)).sum() for f, c in zip(feats, cot)).backward()
    # (a) the kernels, exactly: the same chain rule in float64 on the CPU with the GPU forward's ReLU outputs
    saved = net._eng().vgg_features_train(x.to(dev), "Vgg19")[1]
    exact = _vgg_backward_reference(saved, cot, sd)
    assert rel_linf(xd.grad.cpu().numpy(), exact.numpy()) < 2e-4
    # (b) torch.autograd through the CPU oracle.  A ReLU unit whose pre-activation is within rounding of zero may take the other
    # side there (a handful among ~10^6); each such unit changes the gradient over its receptive field by a few percent, so the
    # comparison is on the relative L2 error and on the bulk of the elements rather than on the maximum
    def close(a, b):
        a, b = a.double(), b.double()
        return float((a - b).norm() / b.norm()) < 2e-2 and float(((a - b).abs() < 2e-3 * b.abs().max()).double().mean()) > 0.9
    assert close(xd.grad.cpu(), xr.grad)
    # only the deepest feature carries a gradient (content loss): the shallower taps contribute nothing
    xd2 = x.to(dev).requires_grad_(True)
    (net.vgg19(xd2).relu4_1 * cot[3].to(dev)).sum().backward()
    assert rel_linf(xd2.grad.cpu().numpy(), _vgg_backward_reference(saved, [None, None, None, cot[3]], sd).numpy()) < 2e-4
    # content + style loss of a "styled" frame against fixed targets, as in train.py:392-399
    target = torch.randn(2, 3, 44, 60, generator=g)
    with torch.no_grad():
        ft = net.vgg19(target.to(dev))
    xd3 = x.to(dev).requires_grad_(True)
    fx = net.vgg19(xd3)
    loss = net.content_loss(fx, ft) + 10.0 * net.style_loss(fx, ft)
    loss.backward()

    def ref_loss(xx):
        fa = stylenet.vgg19_features(xx, sd, top="Vgg19")
        fb = [t.detach() for t in stylenet.vgg19_features(target, sd, top="Vgg19")]
        ms = lambda f: (f.mean((2, 3)), (f.var((2, 3)) + 1e-5).sqrt())
        sl = sum(F.mse_loss(ms(a)[0], ms(b)[0]) + F.mse_loss(ms(a)[1], ms(b)[1]) for a, b in zip(fa, fb))
        return F.mse_loss(fa[3], fb[3]) + 10.0 * sl
    xr3 = x.clone().requires_grad_(True)
    lr = ref_loss(xr3)
    lr.backward()
    assert abs(float(loss) - float(lr)) < 2e-3 * abs(float(lr))
    assert close(xd3.grad.cpu(), xr3.grad)


@pytest.mark.parametrize("shape", [(2, 19, 32, 512, 64), (1, 1, 1, 64, 8), (3, 1, 7, 64, 16), (1, 6, 1, 128, 32), (1, 152, 256, 512, 64)])
def test_conv3x3_output_sum_without_the_convolution(L, dev, shape):
    """rrv_conv3x3_output_sum (FilterPredictor's mean of a convolution from nine border-aware sums of its input) against the
    mean of the convolution itself, including 1-row / 1-column inputs where border rows and corners coincide."""
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, Cin, H, W, generator=g) + 0.25
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    xp = _to_planes(L, x.to(dev))
    xq = _from_planes(L, xp).cpu().double()
    ref = F.conv2d(xq, w.double(), b.double(), padding=1).sum(dim=(0, 2, 3))
    wd, bd = w.to(dev).contiguous(), b.to(dev)
    scratch = torch.empty((9, Cin), dtype=torch.float64, device=dev)
    part = torch.empty((5, Cout), dtype=torch.float64, device=dev)
    L.check(L.lib().rrv_conv3x3_output_sum(L.ptr(xp.hi), L.ptr(xp.lo), N, H, W, Cin, wd.data_ptr(), bd.data_ptr(), Cout, scratch.data_ptr(),
                                           part.data_ptr(), L.stream()))
    part = part.cpu()
    assert torch.all(part[0] == N * H * W)
    scale = float(F.conv2d(xq.abs(), w.double().abs(), b.double().abs(), padding=1).sum(dim=(0, 2, 3)).max())
    assert float((part[1] - ref).abs().max()) < 1e-5 * scale
