"""Dev tool: one convolution through the tcgen05 kernel vs torch CPU, with errors printed."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from rerevst_code_b200 import _lib as L
from rerevst_code_b200.engine import ConvW, Planes, make_epilogue

dev = torch.device("cuda", 0)


def run(N, H, W, Cin, Cout, k, ups, x3=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    hin, win = (H // 2, W // 2) if ups else (H, W)
    x = torch.randn(N, Cin, hin, win, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    cw = ConvW(w.to(dev), b.to(dev), ups=bool(ups))
    xp = Planes(N, hin, win, Cin, x3, dev)
    xd = x.to(dev).contiguous()
    L.check(L.lib().rrv_nchw_to_planes(xd.data_ptr(), N, hin, win, Cin, L.ptr(xp.hi), L.ptr(xp.lo), L.stream()))
    xq = torch.empty_like(xd)
    L.check(L.lib().rrv_planes_to_nchw(L.ptr(xp.hi), L.ptr(xp.lo), N, hin, win, Cin, xq.data_ptr(), L.stream()))
    xq = xq.cpu()
    xin = F.interpolate(xq, scale_factor=2, mode="nearest") if ups else xq
    ref = F.conv2d(xin, w, b, padding=k // 2)
    outs = {}
    for name, impl in (("ffma", 0), ("tc", 1)):
        d = L.Conv()
        d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, Cin, Cout, k, ups
        d.in_hi, d.in_lo = L.ptr(xp.hi), L.ptr(xp.lo)
        d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
        d.ep = make_epilogue(bias=cw.bias)
        d.out_mode = L.OUT_F32_NHWC
        out = torch.full((N, H, W, Cout), float("nan"), dtype=torch.float32, device=dev)
        d.out_f32 = out.data_ptr()
        rc = L.lib().rrv_conv2d(C.byref(d), impl, L.stream())
        if rc != 0:
            print(name, "rc", rc, L.lib().rrv_last_error().decode())
            continue
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(name, "sync failed:", str(e)[:300])
            raise
        got = out.permute(0, 3, 1, 2).cpu()
        err = float((got - ref).abs().max() / ref.abs().max())
        nan = int(torch.isnan(got).sum())
        print(f"{name}: N{N} {H}x{W} {Cin}->{Cout} k{k} ups{ups} x3={x3} rel_linf={err:.3e} nan={nan}")
        outs[name] = got
    return outs, ref


if __name__ == "__main__":
    cases = [(1, 16, 16, 64, 64, 3, 0), (1, 16, 16, 64, 64, 1, 0), (2, 24, 40, 64, 128, 3, 0), (1, 13, 19, 128, 64, 3, 0),
             (1, 16, 32, 256, 128, 3, 1), (1, 24, 48, 256, 512, 3, 0), (1, 26, 38, 128, 64, 3, 1)]
    if len(sys.argv) > 1:
        cases = cases[:int(sys.argv[1])]
    for c in cases:
        run(*c)
