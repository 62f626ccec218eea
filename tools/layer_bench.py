"""Dev tool: time single convolution layers of the 1080p frame under different kernel tunings.
usage: python tools/layer_bench.py  (prints ms per layer per tuning)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rerevst_code_b200 import _lib as L
from rerevst_code_b200.engine import ConvW, Planes, make_epilogue

dev = torch.device("cuda", 0)
LAYERS = {  # name: (H, W, Cin, Cout, k, ups)
    "c64_64": (1216, 2048, 64, 64, 3, 0), "c64_128": (608, 1024, 64, 128, 3, 0), "c128_128": (608, 1024, 128, 128, 3, 0),
    "c256_256": (304, 512, 256, 256, 3, 0), "c256_512": (152, 256, 256, 512, 3, 0), "c512_64": (152, 256, 512, 64, 3, 0),
    "c64_512": (152, 256, 64, 512, 3, 0), "u512_256": (304, 512, 512, 256, 3, 1), "u256_128": (608, 1024, 256, 128, 3, 1),
    "u128_64": (1216, 2048, 128, 64, 3, 1), "s128_64": (608, 1024, 128, 64, 1, 0), "head": (1216, 2048, 64, 3, 3, 0),
    "s512_256": (152, 256, 512, 256, 1, 0), "s256_128": (304, 512, 256, 128, 1, 0),
    "kfup": (152, 256, 64, 512, 3, 0), "kfup3": (152, 256, 64, 512, 3, 0),
    "kfup_32": (152, 256, 32, 512, 3, 0), "kfup_32_3": (152, 256, 32, 512, 3, 0), "kfdown_32": (152, 256, 512, 32, 3, 0),
}


FULL = "--full" in sys.argv
KF = "--kf" in sys.argv        # the KernelFilter up-convolution as the frame runs it: 32 real input channels of 64, + residual planes
                               # (kfup), + Decoder.norm[1] and AdaIN (kfup3)


def bench(name, tunings, iters=5):
    H, W, Cin, Cout, k, ups = LAYERS[name]
    g = torch.Generator().manual_seed(0)
    hin, win = (H // 2, W // 2) if ups else (H, W)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    cw = ConvW(w.to(dev), b.to(dev), ups=bool(ups))
    xp = Planes(1, hin, win, Cin, True, dev)
    xp.hi.normal_()
    xp.lo.zero_()
    d = L.Conv()
    d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = 1, H, W, Cin, Cout, k, ups
    d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
    d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
    if KF:
        d.Cin_used = 32 if Cin == 64 else 0
        res = Planes(1, H, W, Cout, True, dev)
        res.hi.normal_()
        res.lo.zero_()
        if name.endswith("3"):
            tab = torch.stack([torch.zeros(Cout), torch.ones(Cout), torch.full((Cout,), -3.0), torch.full((Cout,), 3.0)]).to(dev).contiguous()
            aff = torch.stack([torch.ones(Cout), torch.zeros(Cout)]).to(dev).contiguous()
            d.ep = make_epilogue(bias=cw.bias, res=res, norm2=tab, affine=aff)
        else:
            d.ep = make_epilogue(bias=cw.bias, res=res)
    elif FULL and Cout % 8 == 0:      # ResidualBlock.conv2 epilogue: norm1, + half-res shortcut, norm2, AdaIN
        tab = torch.stack([torch.zeros(Cout), torch.ones(Cout), torch.full((Cout,), -3.0), torch.full((Cout,), 3.0)]).to(dev).contiguous()
        aff = torch.stack([torch.ones(Cout), torch.zeros(Cout)]).to(dev).contiguous()
        res = Planes(1, H // 2, W // 2, Cout, True, dev)
        res.hi.normal_()
        res.lo.zero_()
        d.ep = make_epilogue(bias=cw.bias, act=2, norm1=tab, res=res, res_shift=1, norm2=tab, affine=aff)
    else:
        d.ep = make_epilogue(bias=cw.bias, act=1)
    if Cout % 8 == 0:
        o = Planes(1, H, W, Cout, True, dev)
        d.out_mode, d.out_hi, d.out_lo = L.OUT_PLANES, L.ptr(o.hi), L.ptr(o.lo)
    else:
        o = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
        d.out_mode, d.out_f32, d.out_C = L.OUT_F32_NCHW, o.data_ptr(), 3
    flops = 2.0 * Cin * Cout * k * k * H * W
    res = []
    for label, (mt, pair, minbn, maxbn, *rest) in tunings.items():
        L.check(L.lib().rrv_tc_tune_merge(rest[0] if rest else 1))
        L.check(L.lib().rrv_tc_tune_pair(pair, minbn))
        L.check(L.lib().rrv_tc_tune(maxbn, mt))
        for _ in range(2):
            L.check(L.lib().rrv_conv2d(C.byref(d), 1, L.stream()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            L.check(L.lib().rrv_conv2d(C.byref(d), 1, L.stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.append(f"{label}={ms:.3f}ms({flops / ms / 1e9:.0f}TF)")
    print(f"{name:10s} " + "  ".join(res), flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:      # the shipped tuning only (A/B of library builds: RRV_LIB_PATH)
        for n in [a for a in sys.argv[1:] if not a.startswith("--")]:
            bench(n, {"default": (2, 1, 64, 256)}, iters=50)
        sys.exit(0)
    if "--small" in sys.argv:
        T = {"default": (2, 1, 64, 256), "mt1": (1, 1, 64, 256), "bn128": (2, 1, 64, 128), "mt1_bn128": (1, 1, 64, 128),
             "mt1_bn64": (1, 1, 64, 64)}
        for n in [a for a in sys.argv[1:] if not a.startswith("--")]:
            bench(n, T, iters=20)
        sys.exit(0)
    T = {"default": (2, 1, 64, 256), "nomerge": (2, 1, 64, 256, 0), "nopair": (2, 0, 64, 256),
         "nopair_nomerge": (2, 0, 64, 256, 0), "pair32": (2, 1, 32, 256)}
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(LAYERS)
    for n in names:
        bench(n, T)
