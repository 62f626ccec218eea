"""Host-side orchestration of the CUDA path: weight repack, activation buffers and the launch
sequences for the per-frame forward, the per-clip pre-pass and the per-style encoder.

All arithmetic on activations happens in ``librerevst_b200.so`` (through the C ABI); torch is
used for device memory, streams and one-off weight algebra (folding the predicted 32x32 dynamic
filters into the neighbouring convolutions, see :meth:`StyleEngine._fold_filter`).

Reference being replaced: ``test/style_network_global.py`` (Encoder :271-281, EncoderStyle
:284-331, Decoder :334-451, TransformerNet :454-501).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict, namedtuple

import torch

from . import _lib as L
from .weights import FILTERS, RES_BLOCKS, VGG_CONVS, VGG_POOL_AFTER, vgg_keys

mean_std = namedtuple("mean_std", ["mean", "std"])
vgg_outputs_super = namedtuple("VggOutputs", ["map", "relu1_1", "relu2_1", "relu3_1", "relu4_1"])

STAT_NAMES = ("norm0", "norm1", "slice4.norm1", "slice4.norm2", "norm2", "slice3.norm1", "slice3.norm2",
              "norm3", "slice2.norm1", "slice2.norm2", "norm4")
INNER_PAD = 32          # channels of the KernelFilter's inner tensor as stored (round 1-2a: zero-padded to 64; the tensor-core path
                        # now reads a 32-channel operand as 64-byte rows)


class Planes:
    """NHWC activation as two 16-bit planes (hi = bf16(v), lo = 16-bit(v - hi)); lo is None in bf16 mode."""
    __slots__ = ("hi", "lo", "N", "H", "W", "C")

    def __init__(self, N, H, W, C, x3, device):
        self.N, self.H, self.W, self.C = N, H, W, C
        self.hi = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=device)
        self.lo = torch.empty((N, H, W, C), dtype=torch.int16, device=device) if x3 else None

    def slice0(self):
        """View of batch item 0."""
        p = Planes.__new__(Planes)
        p.N, p.H, p.W, p.C = 1, self.H, self.W, self.C
        p.hi = self.hi[:1]
        p.lo = self.lo[:1] if self.lo is not None else None
        return p


def _round_up(v, m):
    return (v + m - 1) // m * m


class ConvW:
    """One convolution's weights in the layouts the kernels read, built once at load time
    (SURVEY 8b item vi: state_dict -> kernel layout)."""

    def __init__(self, w_oihw, bias, ups=False, cin_pad=None, cout_pad=None):
        lib = L.lib()
        w = w_oihw.detach().float().contiguous()
        cout, cin, k, _ = w.shape
        cin_p, cout_p = cin_pad or cin, cout_pad or cout
        if (cin_p, cout_p) != (cin, cout):
            wp = torch.zeros((cout_p, cin_p, k, k), dtype=torch.float32, device=w.device)
            wp[:cout, :cin] = w
            w = wp
        self.Cin, self.Cout, self.ksize, self.ups = cin_p, cout_p, k, bool(ups)
        self.Cin_used = cin if cin_p != cin else 0       # zero-padded input channels are not multiplied
        self.bias = None
        if bias is not None:
            self.bias = torch.zeros(cout_p, dtype=torch.float32, device=w.device)
            self.bias[:cout] = bias.detach().float()
        co64 = 8 if cout_p < 8 else _round_up(cout_p, 64)
        self.w_f32 = torch.empty((k * k, cin_p, co64), dtype=torch.float32, device=w.device)
        L.check(lib.rrv_pack_weights_f32(w.data_ptr(), cin_p, cout_p, k, cin_p, co64, self.w_f32.data_ptr(), L.stream()),
                "pack_weights_f32")
        self.w_tc = None
        nbytes = lib.rrv_tc_weight_bytes(cin_p, cout_p, k, int(self.ups))
        if nbytes > 0:
            self.w_tc = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
            L.check(lib.rrv_pack_weights_tc(w.data_ptr(), cin_p, cout_p, k, int(self.ups), self.w_tc.data_ptr(), L.stream()),
                    "pack_weights_tc")
        self._keep = w


class FoldedFilter:
    """The two convolutions of one KernelFilter with its predicted 32x32 matrices folded in, as preallocated tensor-core
    weight blobs that rrv_fold_filter rewrites in place (frame mode: new filters every frame, no allocation, no library GEMM)."""

    class _W:
        __slots__ = ("Cin", "Cout", "ksize", "ups", "Cin_used", "bias", "w_f32", "w_tc")

    def __init__(self, fw, device):
        lib = L.lib()
        self.down, self.up = FoldedFilter._W(), FoldedFilter._W()
        for w, cin, cout, used in ((self.down, 512, INNER_PAD, 0), (self.up, INNER_PAD, 512, 32 if INNER_PAD > 32 else 0)):
            w.Cin, w.Cout, w.ksize, w.ups, w.Cin_used, w.w_f32 = cin, cout, 3, False, used, None
            w.w_tc = torch.empty(lib.rrv_tc_weight_bytes(cin, cout, 3, 0), dtype=torch.uint8, device=device)
        self.down.bias = torch.zeros(INNER_PAD, dtype=torch.float32, device=device)
        self.up.bias = fw["up_b"]
        self._src = (fw["down_w"].contiguous(), fw["down_b"].contiguous(), fw["up_w"].contiguous())

    def fold(self, wf1, wf2):
        dw, db, uw = self._src
        L.check(L.lib().rrv_fold_filter(wf1.data_ptr(), wf2.data_ptr(), dw.data_ptr(), db.data_ptr(), uw.data_ptr(),
                                        self.down.w_tc.data_ptr(), self.down.bias.data_ptr(), self.up.w_tc.data_ptr(), L.stream()),
                "rrv_fold_filter")
        return self.down, self.up


def make_epilogue(bias=None, act=0, norm1=None, res=None, res_shift=0, res_broadcast=False, norm2=None, affine=None):
    e = L.Epilogue()
    e.bias = L.ptr(bias)
    e.act = act
    e.norm1 = L.ptr(norm1)
    if isinstance(res, torch.Tensor):            # fp32 NHWC residual (the 1x1 shortcut of a ResidualBlock)
        e.res_hi, e.res_lo, e.res_f32 = res.data_ptr(), 0, 1
        e.res_shift, e.res_H, e.res_W = res_shift, res.shape[1], res.shape[2]
        e.res_batch_stride = 0 if res_broadcast else res.shape[1] * res.shape[2] * res.shape[3]
    elif res is not None:
        e.res_hi, e.res_lo = L.ptr(res.hi), L.ptr(res.lo)
        e.res_shift, e.res_H, e.res_W = res_shift, res.H, res.W
        e.res_batch_stride = 0 if res_broadcast else res.H * res.W * res.C
    e.norm2 = L.ptr(norm2)
    e.affine = L.ptr(affine)
    e._keep = (bias, norm1, res, norm2, affine)
    return e


class StyleEngine:
    """Everything ``TransformerNet`` needs on one GPU."""

    MAX_PLANS = 6          # captured CUDA graphs kept (LRU by input shape, output kind and lane)

    def __init__(self, device, precision="x3", impl="auto"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("rerevst_b200 runs on CUDA devices only (there is no CPU path); got %s" % device)
        self.lib = L.lib()
        assert precision in ("x3", "bf16")
        self.x3 = precision == "x3"
        self.impl_name = impl
        self.w = None
        self.style = None            # dict: tables + map
        self.F_style = None
        self.stats = {}              # name -> float[4][C] table
        self.filters = {}            # "Filter1" -> (wf1, wf2) fp32 [32,32]
        self.fw = {}                 # "Filter1" -> (ConvW down', ConvW up')
        self.samples = []
        self.q1_sample = None        # Encoder output of the clip's first sampled frame when it lives on another rank (dist.py)
        self.stats_allgather = None  # set by dist.ShardedPrepass: part[5,C] -> parts[G,5,C]
        self._plans = OrderedDict()
        self._sums = set()           # data pointers of one-pass partials whose row 2 still holds sum(x^2) (see _part_done)
        self.profile = None          # list -> (label, start_event, end_event, flops) per conv launch (bench.py)
        self.graph_launches = 0      # kernels launched through CUDA-graph replays (not seen by rrv_launch_count)

    # ------------------------------------------------------------------ weights
    @property
    def impl(self):
        """Kernel family of the per-frame convolutions: the tcgen05 implicit GEMM unless 'ffma' is forced."""
        return L.IMPL_FFMA if self.impl_name == "ffma" else L.IMPL_TCGEN05

    def _impl_for(self, cw):
        if self.impl_name == "ffma":
            return L.IMPL_FFMA
        if cw.w_tc is None:
            if self.impl_name == "tc":
                raise RuntimeError(f"no tensor-core kernel for conv {cw.Cin}->{cw.Cout} k={cw.ksize}")
            return L.IMPL_FFMA
        return L.IMPL_TCGEN05

    def load_weights(self, sd):
        dev = self.device
        g = lambda k: sd[k].detach().to(dev, torch.float32)
        w = {}
        for top in ("Encoder", "EncoderStyle", "Vgg19"):
            if top == "Vgg19" and vgg_keys(top)[0][0] not in sd:       # deleted after the first style, like the reference
                continue
            convs = []
            for j, (wk, bk) in enumerate(vgg_keys(top)):
                if j == 0:
                    convs.append((g(wk).contiguous(), g(bk).contiguous()))     # conv1_1: rrv_first_layer
                else:
                    convs.append(ConvW(g(wk), g(bk)))
            w[top] = convs
        for name, cin, cout in RES_BLOCKS:
            p = f"Decoder.{name}."
            w[name] = dict(conv1=ConvW(g(p + "conv1.weight"), g(p + "conv1.bias"), ups=True),
                           conv2=ConvW(g(p + "conv2.weight"), g(p + "conv2.bias")),
                           short=ConvW(g(p + "conv_shortcut.weight"), None))
        w["slice1"] = ConvW(g("Decoder.slice1.weight"), g("Decoder.slice1.bias"))
        for f in FILTERS:
            p = f"Decoder.{f}."
            pred_w = torch.cat([g(p + "F1.down_sample.0.weight"), g(p + "F2.down_sample.0.weight")], 0)
            pred_b = torch.cat([g(p + "F1.down_sample.0.bias"), g(p + "F2.down_sample.0.bias")], 0)
            w[f] = dict(down_w=g(p + "down_sample.0.weight"), down_b=g(p + "down_sample.0.bias"),
                        up_w=g(p + "upsample.0.weight"), up_b=g(p + "upsample.0.bias"),
                        pred=ConvW(pred_w, pred_b),            # F1 | F2 predictor convs stacked: 512 -> 64
                        fc=[(g(p + f"{q}.FC.weight").contiguous(), g(p + f"{q}.FC.bias").contiguous()) for q in ("F1", "F2")])
        self.w = w
        self.fw = {}
        self._plans = OrderedDict()
        if self.filters:                       # re-fold cached filters against the new weights
            for f, (a, b) in self.filters.items():
                self._fold_filter(f, a, b)

    # ------------------------------------------------------------------ thin wrappers over the C ABI
    def _conv(self, cw, x, ep, out_mode=L.OUT_PLANES, out=None, N=None, out_C=0, pool=False, crop=None, terms=0, stats=None,
              stats_minmax=False):
        N = x.N if N is None else N
        H, W = (x.H * 2, x.W * 2) if cw.ups else (x.H, x.W)
        if pool and (self._impl_for(cw) != L.IMPL_TCGEN05 or cw.Cout % 32 != 0 or H < 2 or W < 2):
            return self._pool(self._conv(cw, x, ep, out_mode, None, N, out_C))       # FFMA bring-up path: separate pool kernel
        assert x.C == cw.Cin, (x.C, cw.Cin)
        d = L.Conv()
        d.N, d.H, d.W, d.Cin, d.Cout, d.ksize, d.ups = N, H, W, cw.Cin, cw.Cout, cw.ksize, int(cw.ups)
        d.in_hi, d.in_lo = L.ptr(x.hi), L.ptr(x.lo)
        d.w_f32, d.w_tc = L.ptr(cw.w_f32), L.ptr(cw.w_tc)
        d.ep = ep
        d.out_mode = out_mode
        d.pool = int(pool)
        d.Cin_used = cw.Cin_used
        d.terms = terms
        if stats is not None:
            d.stats, d.stats_minmax = stats.data_ptr(), int(stats_minmax)
        if out_mode in (L.OUT_BGR_F32, L.OUT_BGR_U8):      # the RGB head writes the post-processed, cropped HWC BGR frame
            y0, x0, h, w = crop if crop is not None else (0, 0, H, W)
            dt = torch.float32 if out_mode == L.OUT_BGR_F32 else torch.uint8
            out = out if out is not None else torch.empty((N, h, w, 3), dtype=dt, device=self.device)
            if tuple(out.shape) != (N, h, w, 3) or out.dtype != dt or not out.is_contiguous():
                raise ValueError(f"frame output is {tuple(out.shape)} {out.dtype}, expected contiguous {dt} {(N, h, w, 3)}")
            d.out_img = out.data_ptr()
            d.crop_y0, d.crop_x0, d.crop_h, d.crop_w = y0, x0, h, w
        elif out_mode == L.OUT_PLANES:
            out = out or Planes(N, H >> int(pool), W >> int(pool), cw.Cout, self.x3, self.device)
            if (out.N, out.H, out.W, out.C) != (N, H >> int(pool), W >> int(pool), cw.Cout):
                raise ValueError(f"conv output planes are {(out.N, out.H, out.W, out.C)}, expected {(N, H >> int(pool), W >> int(pool), cw.Cout)}")
            d.out_hi, d.out_lo = L.ptr(out.hi), L.ptr(out.lo)
        elif out_mode == L.OUT_F32_NHWC:
            out = out if out is not None else torch.empty((N, H, W, cw.Cout), dtype=torch.float32, device=self.device)
            self._check_out(out, (N, H, W, cw.Cout))
            d.out_f32 = out.data_ptr()
        else:
            out = out if out is not None else torch.empty((N, out_C, H, W), dtype=torch.float32, device=self.device)
            self._check_out(out, (N, out_C, H, W))
            d.out_f32, d.out_C = out.data_ptr(), out_C
        if self.profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        L.check(self.lib.rrv_conv2d(C.byref(d), self._impl_for(cw), L.stream()), "rrv_conv2d")
        if self.profile is not None:
            e1.record()
            hin, win = x.H, x.W
            # executed tensor work: 4 of 9 taps for the nearest-x2 layers, k-slices of 16 real input channels, Cout padded to 16,
            # three MMAs per k-slice in x3 mode
            taps = 4 if cw.ups else cw.ksize * cw.ksize
            k_exec = -(-(cw.Cin_used or cw.Cin) // 16) * 16
            executed = 2.0 * k_exec * (-(-cw.Cout // 16) * 16) * taps * N * H * W * (3 if x.lo is not None else 1)
            self.profile.append((f"conv{cw.ksize}x{cw.ksize}{'u' if cw.ups else ''} {cw.Cin}->{cw.Cout} @{H}x{W}", e0, e1,
                                 2.0 * cw.Cin * cw.Cout * cw.ksize * cw.ksize * N * H * W, executed))
        return out

    @staticmethod
    def _check_out(out, shape):
        """A caller-supplied fp32 output must be exactly the tensor the kernel writes (contiguous, right shape): the kernels
        compute their own strides from the convolution's size."""
        if tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"conv output tensor is {tuple(out.shape)} {out.dtype} (contiguous={out.is_contiguous()}), "
                             f"expected contiguous float32 {tuple(shape)}")

    @staticmethod
    def output_size(H, W):
        """Spatial size of TransformerNet.forward's result for an H x W input: three floor max-pools (vgg19.features[4|9|18])
        then three nearest x2 upsamples -- (H // 8) * 8, e.g. 436 x 1024 -> 432 x 1024, like the reference."""
        return (H // 8) * 8, (W // 8) * 8

    def _pointwise(self, x_f32, ep, N=None, broadcast=False, to_f32=False, stats=None, stats_minmax=False):
        n_in, H, W, Cc = x_f32.shape
        N = n_in if N is None else N
        bs = 0 if broadcast else H * W * Cc
        if to_f32:
            out = torch.empty((N, H, W, Cc), dtype=torch.float32, device=self.device)
            args = (x_f32.data_ptr(), bs, N, H, W, Cc, C.byref(ep), L.OUT_F32_NHWC, 0, 0, out.data_ptr())
        else:
            out = Planes(N, H, W, Cc, self.x3, self.device)
            args = (x_f32.data_ptr(), bs, N, H, W, Cc, C.byref(ep), L.OUT_PLANES, L.ptr(out.hi), L.ptr(out.lo), 0)
        if stats is not None:
            L.check(self.lib.rrv_pointwise_stats(*args, stats.data_ptr(), int(stats_minmax), L.stream()), "rrv_pointwise_stats")
        else:
            L.check(self.lib.rrv_pointwise(*args, L.stream()), "rrv_pointwise")
        return out

    # ---- one-pass statistics: the kernel that writes a tensor also accumulates its per-channel sums (rrv_conv.stats) ----
    @property
    def fused_stats(self):
        return self.impl_name != "ffma"

    def _part_new(self, Cc, npix):
        part = torch.empty((5, Cc), dtype=torch.float64, device=self.device)
        L.check(self.lib.rrv_stats_init(part.data_ptr(), Cc, float(npix), L.stream()), "rrv_stats_init")
        return part

    def _part_done(self, part):
        """{n, sum, sum of squares, min, max} -> {n, sum, M2, min, max}, merged over the ranks of a sharded pre-pass.  In one
        process the conversion is left to rrv_stats_finalize (kind | 16): one small launch less per statistic point."""
        if self.stats_allgather is None:
            self._sums.add(part.data_ptr())
            return part
        Cc = part.shape[1]
        L.check(self.lib.rrv_stats_sums_to_m2(part.data_ptr(), Cc, L.stream()), "rrv_stats_sums_to_m2")
        return self._merge_ranks(part)

    def _merge_ranks(self, part):
        if self.stats_allgather is None:
            return part
        Cc = part.shape[1]
        parts = self.stats_allgather(part).contiguous()              # frame-parallel pre-pass (dist.py)
        merged = torch.empty_like(part)
        L.check(self.lib.rrv_stats_merge(parts.data_ptr(), parts.shape[0], Cc, merged.data_ptr(), L.stream()), "rrv_stats_merge")
        return merged

    def _pred_content_part(self, cw, x):
        """Partial {count, sum} of conv3x3(x) + b over all output pixels WITHOUT running the convolution: FilterPredictor only
        needs its mean (style_network_global.py:163-167), and that follows from nine border-aware channel sums of x
        (rrv_conv3x3_output_sum).  Merged over the ranks of a sharded pre-pass like every other partial."""
        scratch = torch.empty((9, cw.Cin), dtype=torch.float64, device=self.device)
        part = torch.empty((5, cw.Cout), dtype=torch.float64, device=self.device)
        L.check(self.lib.rrv_conv3x3_output_sum(L.ptr(x.hi), L.ptr(x.lo), x.N, x.H, x.W, cw.Cin, cw._keep.data_ptr(), L.ptr(cw.bias),
                                                cw.Cout, scratch.data_ptr(), part.data_ptr(), L.stream()), "rrv_conv3x3_output_sum")
        return self._merge_ranks(part)

    def _conv_stats(self, cw, x, ep, minmax, N=None):
        """fp32 NHWC convolution output + the finished partial statistics of that output (one pass on the tensor-core path)."""
        if not self.fused_stats or self._impl_for(cw) != L.IMPL_TCGEN05:
            out = self._conv(cw, x, ep, L.OUT_F32_NHWC, N=N)
            return out, self._stats_part(out)
        n = x.N if N is None else N
        H, W = (x.H * 2, x.W * 2) if cw.ups else (x.H, x.W)
        part = self._part_new(cw.Cout, n * H * W)
        out = self._conv(cw, x, ep, L.OUT_F32_NHWC, N=N, stats=part, stats_minmax=minmax)
        return out, self._part_done(part)

    def _pointwise_stats(self, x_f32, ep, minmax, **kw):
        """pointwise pass + the finished partial statistics of what it wrote."""
        n_in, H, W, Cc = x_f32.shape
        n = kw.get("N") or n_in
        if not self.fused_stats:
            out = self._pointwise(x_f32, ep, to_f32=True, **kw)
            return out, self._stats_part(out)
        part = self._part_new(Cc, n * H * W)
        out = self._pointwise(x_f32, ep, to_f32=True, stats=part, stats_minmax=minmax, **kw)
        return out, self._part_done(part)

    def _pool(self, x):
        out = Planes(x.N, x.H // 2, x.W // 2, x.C, self.x3, self.device)
        L.check(self.lib.rrv_maxpool2x2(L.ptr(x.hi), L.ptr(x.lo), x.N, x.H, x.W, x.C, L.ptr(out.hi), L.ptr(out.lo),
                                        L.stream()), "rrv_maxpool2x2")
        return out

    def _stats_part(self, x_f32):
        Cc = x_f32.shape[-1]
        part = torch.empty((5, Cc), dtype=torch.float64, device=self.device)
        L.check(self.lib.rrv_channel_stats(x_f32.data_ptr(), x_f32.numel() // Cc, Cc, part.data_ptr(), L.stream()),
                "rrv_channel_stats")
        return self._merge_ranks(part)

    def _finalize(self, part, kind, eps):
        Cc = part.shape[1]
        rows = {0: 4, 1: 2, 2: 1, 3: 4}[kind]
        out = torch.empty((rows, Cc), dtype=torch.float32, device=self.device)
        if part.data_ptr() in self._sums:              # a one-pass partial: row 2 is still the sum of squares
            self._sums.discard(part.data_ptr())
            kind |= 16
        L.check(self.lib.rrv_stats_finalize(part.data_ptr(), Cc, kind, eps, out.data_ptr(), L.stream()), "rrv_stats_finalize")
        return out

    def _saved_stat(self, x_f32):
        """InstanceNorm.compute (style_network_global.py:59-77) -> float[4][C] table."""
        return self._finalize(self._stats_part(x_f32), 0, 1e-8)

    # ------------------------------------------------------------------ VGG stacks
    def _vgg(self, top, src, kind, gray, N, H, W, mode, norm0=None):
        """conv1_1 .. conv4_1.  mode 'content': planes of norm0(relu4_1); 'raw': fp32 NHWC relu4_1;
        'style': dict of fp32 NHWC taps at relu1_1 / 2_1 / 3_1 / 4_1."""
        convs = self.w[top]
        w0, b0 = convs[0]
        x = Planes(N, H, W, 64, self.x3, self.device)
        taps = {}
        f32 = torch.empty((N, H, W, 64), dtype=torch.float32, device=self.device) if mode == "style" else None
        L.check(self.lib.rrv_first_layer(src.data_ptr(), kind, int(gray), N, H, W, w0.data_ptr(), b0.data_ptr(),
                                         L.ptr(x.hi), L.ptr(x.lo), L.ptr(f32), L.stream()), "rrv_first_layer")
        if mode == "style":
            taps[0] = f32
        ident = make_epilogue()
        for (idx, _, _), cw in zip(VGG_CONVS[1:], convs[1:]):
            last = idx == 19
            ep = make_epilogue(bias=cw.bias, act=1)
            if mode == "style" and idx in (5, 10, 19):
                taps[idx] = self._conv(cw, x, ep, L.OUT_F32_NHWC)
                if not last:
                    x = self._pointwise(taps[idx], ident)
            elif last and mode == "raw":
                return self._conv(cw, x, ep, L.OUT_F32_NHWC)
            elif last and mode == "raw+stats":
                return self._conv_stats(cw, x, ep, False)
            elif last:
                return self._conv(cw, x, make_epilogue(bias=cw.bias, act=1, norm1=norm0))
            else:
                x = self._conv(cw, x, ep, pool=idx in VGG_POOL_AFTER)     # MaxPool2d fused into the epilogue
        return taps

    # ------------------------------------------------------------------ style (once per style image)
    @torch.no_grad()
    def generate_style_features(self, style, kind=0):
        """EncoderStyle.forward + cal_mean_std (style_network_global.py:304-331).
        style: fp32 NCHW [1,3,h,w] normalised (kind 0) or uint8 HWC BGR [1,h,w,3] (kind 1)."""
        if kind == 0:
            N, _, H, W = style.shape
        else:
            N, H, W, _ = style.shape
        assert N == 1
        style = style.contiguous()
        taps = self._vgg("EncoderStyle", style, kind, False, N, H, W, "style")
        tabs = {}
        for lvl, idx in (("relu1_1", 0), ("relu2_1", 5), ("relu3_1", 10), ("relu4_1", 19)):
            x = taps[idx]
            Cc = x.shape[-1]
            part = torch.empty((5, Cc), dtype=torch.float64, device=self.device)
            L.check(self.lib.rrv_channel_stats(x.data_ptr(), x.numel() // Cc, Cc, part.data_ptr(), L.stream()), "stats")
            tabs[lvl] = self._finalize(part, 1, 1e-5)          # float[2][C] = {std, mean} = AdaIN {scale, shift}
        smap = taps[19]                                       # [1, h/8, w/8, 512] fp32
        # normalized_style = (map - mean) / std  (:371, :397), carried as planes for the predictor convs
        sc, sh = tabs["relu4_1"][0], tabs["relu4_1"][1]
        ntab = torch.stack([sh, 1.0 / sc, torch.full_like(sc, float("-inf")), torch.full_like(sc, float("inf"))]).contiguous()
        nstyle = self._pointwise(smap, make_epilogue(norm1=ntab))
        self.style = dict(tabs=tabs, map=smap, nstyle=nstyle)
        ms = {k: mean_std(v[1].view(1, -1, 1, 1), v[0].view(1, -1, 1, 1)) for k, v in tabs.items()}
        self.F_style = vgg_outputs_super(smap.permute(0, 3, 1, 2), ms["relu1_1"], ms["relu2_1"], ms["relu3_1"], ms["relu4_1"])
        self._plans = OrderedDict()

    # ------------------------------------------------------------------ pre-pass
    def clean(self):
        self.samples = []
        self.q1_sample = None
        self.stats = {}
        self.filters = {}
        self.fw = {}
        self._plans = OrderedDict()

    @torch.no_grad()
    def add(self, patch, kind=0):
        """TransformerNet.add (:471-475): Encoder(RGB2Gray(patch)) kept for compute()."""
        if kind == 0:
            N, _, H, W = patch.shape
        else:
            N, H, W, _ = patch.shape
        self.samples.append(self._vgg("Encoder", patch.contiguous(), kind, True, N, H, W, "raw"))

    @torch.no_grad()
    def add_q1(self, patch, kind=0):
        """Frame-parallel pre-pass: the clip's FIRST sampled frame, encoded on a rank that does not own
        it.  Quirk Q1 (KernelFilter.compute, :223-230) filters only sample 0 and broadcasts its residual to
        every sample, so every rank carries sample 0 through norm[0] -> Filter1..3; it does not enter this
        rank's statistics."""
        if kind == 0:
            N, _, H, W = patch.shape
        else:
            N, H, W, _ = patch.shape
        self.q1_sample = self._vgg("Encoder", patch.contiguous(), kind, True, N, H, W, "raw")

    def _fold_filter(self, f, wf1, wf2):
        """Fold the two predicted 32x32 matrices of a KernelFilter (apply_filter, :194-217) into its
        down_sample / upsample convolutions: Wf1 . conv_down(x) == conv_{Wf1.Wdown}(x) and
        conv_up(Wf2 . t) == conv_{Wup.Wf2}(t).  One-off per clip, fp32."""
        fw = self.w[f]
        if self.impl_name != "ffma":                      # on the device, straight into tensor-core blobs (rrv_fold_filter)
            ff = FoldedFilter(fw, self.device)
            self.fw[f] = ff.fold(wf1.contiguous(), wf2.contiguous())
            self._fold_keep = getattr(self, "_fold_keep", {})
            self._fold_keep[f] = ff
        else:                                             # FFMA bring-up path: fp32 weight layout through the generic repack
            dw = torch.matmul(wf1, fw["down_w"].reshape(32, -1)).reshape(32, 512, 3, 3)
            db = torch.mv(wf1, fw["down_b"])
            uw = torch.einsum("ojyx,ji->oiyx", fw["up_w"], wf2).contiguous()
            self.fw[f] = (ConvW(dw, db, cout_pad=INNER_PAD), ConvW(uw, fw["up_b"], cin_pad=INNER_PAD))
        self.filters[f] = (wf1, wf2)

    def _predict_filters(self, f, content):
        """FilterPredictor.compute for F1 and F2 of one KernelFilter (:161-172): both predictor convs
        run as one 512 -> 64 convolution; spatial+batch mean; Linear(64 -> 1024) each."""
        fw = self.w[f]
        ep = make_epilogue(bias=fw["pred"].bias)
        c_mean = self._finalize(self._pred_content_part(fw["pred"], content), 2, 0.0)[0]
        s = self._conv(fw["pred"], self.style["nstyle"], ep, L.OUT_F32_NHWC)
        part = torch.empty((5, 64), dtype=torch.float64, device=self.device)
        L.check(self.lib.rrv_channel_stats(s.data_ptr(), s.numel() // 64, 64, part.data_ptr(), L.stream()), "stats")
        s_mean = self._finalize(part, 2, 0.0)[0]
        out = []
        for j, (fcw, fcb) in enumerate(fw["fc"]):
            wf = torch.empty((32, 32), dtype=torch.float32, device=self.device)
            L.check(self.lib.rrv_filter_fc(fcw.data_ptr(), fcb.data_ptr(), c_mean[32 * j:].data_ptr(),
                                           s_mean[32 * j:].data_ptr(), wf.data_ptr(), L.stream()), "rrv_filter_fc")
            out.append(wf)
        return out

    @torch.no_grad()
    def compute(self, keep_samples=False):
        """Decoder.compute (:425-439): 11 global statistic tables + 6 dynamic filters, stage by stage
        over all sampled frames.  Quirk Q1 is reproduced: only sample 0 goes through the dynamic
        filters and its residual is broadcast to every sample.  keep_samples: the multi-style model runs
        this once per style on the same sampled frames."""
        if self.style is None:
            raise RuntimeError("compute() before generate_style_features()")
        if not self.samples:
            raise RuntimeError("compute() without add(): no sampled frames")
        x = torch.cat(self.samples, 0) if len(self.samples) > 1 else self.samples[0]
        N = x.shape[0]
        st = self.stats
        tabs = self.style["tabs"]
        st["norm0"] = self._saved_stat(x)
        h = self._pointwise(x, make_epilogue(norm1=st["norm0"]))            # planes [N, h, w, 512]
        # sample 0 of the CLIP (Q1): local sample 0 on the rank that owns it, q1_sample elsewhere
        h0 = self._pointwise(self.q1_sample, make_epilogue(norm1=st["norm0"])) if self.q1_sample is not None else None
        del x
        for i, f in enumerate(FILTERS):
            wf1, wf2 = self._predict_filters(f, h)
            self._fold_filter(f, wf1, wf2)
            down, up = self.fw[f]
            t = self._conv(down, h0 if h0 is not None else h.slice0(), make_epilogue(bias=down.bias, act=2))
            u0 = self._conv(up, t, make_epilogue(bias=up.bias), L.OUT_F32_NHWC)   # [1,h,w,512]
            if h0 is not None and i < 2:
                h0 = self._pointwise(u0, make_epilogue(res=h0))
            if i < 2:
                h = self._pointwise(u0, make_epilogue(res=h), N=N, broadcast=True)
            else:
                r, rpart = self._pointwise_stats(u0, make_epilogue(res=h), True, N=N, broadcast=True)
        levels = (("norm1", "relu4_1", "slice4"), ("norm2", "relu3_1", "slice3"),
                  ("norm3", "relu2_1", "slice2"), ("norm4", "relu1_1", None))
        # every statistic comes out of the kernel that writes the tensor (rrv_conv.stats / rrv_pointwise_stats): no extra passes
        for nname, lvl, block in levels:
            st[nname] = self._finalize(rpart, 0, 1e-8)
            if block is None:
                break
            h = self._pointwise(r, make_epilogue(norm1=st[nname], affine=tabs[lvl]))
            del r
            bw = self.w[block]
            s = self._conv(bw["short"], h, make_epilogue())
            r1, part = self._conv_stats(bw["conv1"], h, make_epilogue(bias=bw["conv1"].bias, act=2), True)
            st[block + ".norm1"] = self._finalize(part, 0, 1e-8)
            h = self._pointwise(r1, make_epilogue(norm1=st[block + ".norm1"]))
            del r1
            r2, part = self._conv_stats(bw["conv2"], h, make_epilogue(bias=bw["conv2"].bias, act=2), True)
            st[block + ".norm2"] = self._finalize(part, 0, 1e-8)
            r, rpart = self._pointwise_stats(r2, make_epilogue(norm1=st[block + ".norm2"], res=s, res_shift=1), True)
            del r2
        if not keep_samples:
            self.samples = []
            self.q1_sample = None
        self._plans = OrderedDict()

    # ------------------------------------------------------------------ per-frame forward
    def _require_ready(self):
        if self.style is None:
            raise RuntimeError("forward() before generate_style_features()")
        if len(self.stats) != len(STAT_NAMES) or len(self.fw) != 3:
            raise RuntimeError("forward() before compute(): the global statistics are not available "
                               "(the reference fails here with AttributeError: 'NoneType' has no attribute 'expand')")

    @torch.no_grad()
    def forward(self, frame, kind=0, out=None, post=None):
        """TransformerNet.forward (:499-501) for a batch of independent frames.
        frame: fp32 NCHW normalised (kind 0) or uint8 NHWC BGR (kind 1).  Returns fp32 NCHW, or with
        ``post = ("f32" | "u8", (y0, x0, h, w) | None)`` the finished frame of framework.transfer: transform_back_image +
        tensor2numpy (test/framework.py:39-49) and the crop of generate_real_video.py:167 run in the RGB head's epilogue and the
        result is [N, h, w, 3] BGR in [0, 255] (fp32, or uint8 as cv2.imwrite would store it)."""
        self._require_ready()
        if kind == 0:
            N, _, H, W = frame.shape
        else:
            N, H, W, _ = frame.shape
        h = self._vgg("Encoder", frame.contiguous(), kind, True, N, H, W, "content", norm0=self.stats["norm0"])
        return self._decode(h, out, post)

    @torch.no_grad()
    def encode(self, frame, kind=0):
        """Encoder(RGB2Gray(frame)) (:280-281, :487-497) -> relu4_1 as the fp32 NCHW view of an NHWC tensor
        (generate_content_features of the multi-style model)."""
        if kind == 0:
            N, _, H, W = frame.shape
        else:
            N, H, W, _ = frame.shape
        return self._vgg("Encoder", frame.contiguous(), kind, True, N, H, W, "raw").permute(0, 3, 1, 2)

    @torch.no_grad()
    def decode_features(self, f_content, out=None):
        """Decoder.forward (:441-451) on encoder features [N,512,h,w] (fp32, ideally the view returned by encode())."""
        self._require_ready()
        x = f_content.permute(0, 2, 3, 1).contiguous()
        return self._decode(self._pointwise(x, make_epilogue(norm1=self.stats["norm0"])), out)

    def _head(self, h, out=None, post=None):
        """Decoder.slice1 (:341, :450); ``post`` as in forward()."""
        head = self.w["slice1"]
        if post is None:
            return self._conv(head, h, make_epilogue(bias=head.bias), L.OUT_F32_NCHW, out=out, out_C=3)
        mode, crop = post
        if mode not in ("f32", "u8"):
            raise ValueError("post mode must be 'f32' or 'u8'")
        if self._impl_for(head) != L.IMPL_TCGEN05:      # bring-up path: separate postprocess kernel
            y = self._conv(head, h, make_epilogue(bias=head.bias), L.OUT_F32_NCHW, out_C=3)
            return self.postprocess(y, crop, mode, out)
        return self._conv(head, h, make_epilogue(bias=head.bias), L.OUT_BGR_U8 if mode == "u8" else L.OUT_BGR_F32, out=out, crop=crop)

    def postprocess(self, y, crop=None, mode="f32", out=None):
        """transform_back_image + tensor2numpy (test/framework.py:39-49) + crop on an fp32 NCHW network output."""
        N, _, H, W = y.shape
        y0, x0, h, w = crop if crop is not None else (0, 0, H, W)
        dt = torch.uint8 if mode == "u8" else torch.float32
        out = out if out is not None else torch.empty((N, h, w, 3), dtype=dt, device=self.device)
        fn = self.lib.rrv_postprocess_bgr_u8 if mode == "u8" else self.lib.rrv_postprocess_bgr
        L.check(fn(y.contiguous().data_ptr(), N, H, W, y0, x0, h, w, out.data_ptr(), L.stream()), "rrv_postprocess_bgr")
        return out

    def _decode(self, h, out=None, post=None):
        """norm[0](relu4_1) planes -> Filter1..3 -> AdaIN -> slice4..2 -> slice1 (Decoder.forward :441-451)."""
        st, tabs = self.stats, self.style["tabs"]
        for i, f in enumerate(FILTERS):
            down, up = self.fw[f]
            t = self._conv(down, h, make_epilogue(bias=down.bias, act=2))
            if i < 2:
                h = self._conv(up, t, make_epilogue(bias=up.bias, res=h))
            else:   # + Decoder.norm[1] and AdaIN(relu4_1) (:443)
                h = self._conv(up, t, make_epilogue(bias=up.bias, res=h, norm2=st["norm1"], affine=tabs["relu4_1"]))
        for block, nxt, lvl in (("slice4", "norm2", "relu3_1"), ("slice3", "norm3", "relu2_1"), ("slice2", "norm4", "relu1_1")):
            bw = self.w[block]
            s = self._conv(bw["short"], h, make_epilogue(), L.OUT_F32_NHWC)   # 1x1 shortcut at low resolution, kept in fp32
            y = self._conv(bw["conv1"], h, make_epilogue(bias=bw["conv1"].bias, act=2, norm1=st[block + ".norm1"]))
            h = self._conv(bw["conv2"], y, make_epilogue(bias=bw["conv2"].bias, act=2, norm1=st[block + ".norm2"],
                                                        res=s, res_shift=1, norm2=st[nxt], affine=tabs[lvl]))
        return self._head(h, out, post)

    @torch.no_grad()
    def forward_graphed(self, frame, kind=0, post=None, lane=0):
        """forward() replayed from a CUDA graph: the ~30 launches of a frame (with their TMA descriptors baked
        in) are captured once per input shape and clip state, then each call is one D2D copy of the frame into
        the graph's static input plus one graph launch.  Returns the graph's static output tensor (valid until
        the next call).  Any change of weights, style or clip statistics drops the captured graphs.

        ``lane``: frames are independent, so a caller may keep two of them in flight on two streams (lane_streams()): each
        lane owns its graph, static input / output and activation pool.  Every layer's persistent grid ends with a partly
        filled last round (conv4_1: 160 tiles on 74 CTA pairs = 3 rounds for 2.05 rounds of work) and a launch gap; the other
        lane's kernels fill those SMs."""
        self._require_ready()
        pdl = self.__dict__.get("_lanes", 1) == 1      # several frames in flight: no early launch of the next layer (rrv_tc_tune_pdl)
        key = ("graph", kind, tuple(frame.shape), frame.dtype, post, lane, pdl)
        plan = self._plans.get(key)
        if plan is None:
            if kind == 0:
                N, _, H, W = frame.shape
            else:
                N, H, W, _ = frame.shape
            g_in = torch.empty_like(frame, memory_format=torch.contiguous_format)
            g_in.copy_(frame)
            oh, ow = self.output_size(H, W)
            if post is None:
                g_out = torch.empty((N, 3, oh, ow), dtype=torch.float32, device=self.device)
            else:
                crop = post[1] if post[1] is not None else (0, 0, oh, ow)
                g_out = torch.empty((N, crop[2], crop[3], 3), dtype=torch.uint8 if post[0] == "u8" else torch.float32, device=self.device)
            n0 = self.lib.rrv_launch_count()
            self.forward(g_in, kind, out=g_out, post=post)          # eager warm-up: function attributes, allocator pools
            launches = self.lib.rrv_launch_count() - n0
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            L.check(self.lib.rrv_tc_tune_pdl(1 if pdl else 0), "rrv_tc_tune_pdl")       # baked into the captured launches
            try:
                with torch.cuda.graph(graph):
                    self.forward(g_in, kind, out=g_out, post=post)
            finally:
                self.lib.rrv_tc_tune_pdl(1)
            plan = (graph, g_in, g_out, launches)
            self._plans[key] = plan
            while len(self._plans) > self.MAX_PLANS:        # each plan pins a private pool with a frame's activations
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        graph, g_in, g_out, launches = plan
        g_in.copy_(frame, non_blocking=True)
        graph.replay()
        self.graph_launches += launches
        return g_out

    def lane_streams(self, n):
        """Streams for ``n`` frames in flight: [the current stream, side streams ...] (forward_graphed(..., lane=i) is called with
        stream i current).  The side streams first wait for everything queued on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        self._lanes = max(1, int(n))
        side = self.__dict__.setdefault("_lane_side", [])
        while len(side) < n - 1:
            side.append(torch.cuda.Stream(self.device))
        for st in side[:n - 1]:
            st.wait_stream(cur)
        return [cur] + side[:n - 1]

    # ------------------------------------------------------------------ frame mode (use_Global=False) and training-side VGG
    def _frame_norm(self, x_f32):
        """Frame-mode InstanceNorm (style_network_frame.py:39-43): per-frame mean / rsqrt(biased var + 1e-8), no clamp.
        Returns the float[4][C] table {mean, rstd, -inf, +inf} of ONE sample."""
        Cc = x_f32.shape[-1]
        part = torch.empty((5, Cc), dtype=torch.float64, device=self.device)
        L.check(self.lib.rrv_channel_stats(x_f32.data_ptr(), x_f32.numel() // Cc, Cc, part.data_ptr(), L.stream()),
                "rrv_channel_stats")
        return self._finalize(part, 3, 1e-8)

    def _style_pred_means(self, f):
        """mean_hw(F?.down_sample(normalized_style)) of both predictors of a KernelFilter: frame-invariant, cached per style."""
        cache = self.style.setdefault("pred_means", {})
        if f not in cache:
            fw = self.w[f]
            s = self._conv(fw["pred"], self.style["nstyle"], make_epilogue(bias=fw["pred"].bias), L.OUT_F32_NHWC)
            part = torch.empty((5, 64), dtype=torch.float64, device=self.device)
            L.check(self.lib.rrv_channel_stats(s.data_ptr(), s.numel() // 64, 64, part.data_ptr(), L.stream()), "stats")
            cache[f] = self._finalize(part, 2, 0.0)[0]
        return cache[f]

    @torch.no_grad()
    def forward_frame(self, frame, kind=0, gray=True, out=None):
        """TransformerNet.forward of test/style_network_frame.py:392-394 (Decoder.forward :341-358): every InstanceNorm takes
        its statistics from the frame itself, the six dynamic filters are predicted per frame (FilterPredictor.forward
        :53-62), AdaIN(relu4_1) follows the filters directly (:339).  gray=False is ``validation`` of
        train/style_networks.py:556-559.  A batch is processed sample by sample (the statistics are per sample)."""
        if self.style is None:
            raise RuntimeError("forward() before generate_style_features()")
        if kind == 0:
            N, _, H, W = frame.shape
        else:
            N, H, W, _ = frame.shape
        frame = frame.contiguous()
        if out is None:
            out = torch.empty((N, 3) + self.output_size(H, W), dtype=torch.float32, device=self.device)
        else:
            self._check_out(out, (N, 3) + self.output_size(H, W))
        tabs = self.style["tabs"]
        for i in range(N):
            x, part = self._vgg("Encoder", frame[i:i + 1], kind, gray, 1, H, W, "raw+stats")   # fp32 NHWC relu4_1 + its statistics
            h = self._pointwise(x, make_epilogue(norm1=self._finalize(part, 3, 1e-8)))
            for j, f in enumerate(FILTERS):
                fw = self.w[f]
                ep = make_epilogue(bias=fw["pred"].bias)
                c_mean = self._finalize(self._pred_content_part(fw["pred"], h), 2, 0.0)[0]
                s_mean = self._style_pred_means(f)
                wf = []
                for q, (fcw, fcb) in enumerate(fw["fc"]):
                    m = torch.empty((32, 32), dtype=torch.float32, device=self.device)
                    L.check(self.lib.rrv_filter_fc(fcw.data_ptr(), fcb.data_ptr(), c_mean[32 * q:].data_ptr(),
                                                   s_mean[32 * q:].data_ptr(), m.data_ptr(), L.stream()), "rrv_filter_fc")
                    wf.append(m)
                down, up = self._folded(f, wf[0], wf[1])
                t = self._conv(down, h, make_epilogue(bias=down.bias, act=2))
                if j < 2:
                    h = self._conv(up, t, make_epilogue(bias=up.bias, res=h))
                else:       # results * style_std + style_mean (:339), no second norm in this mode
                    h = self._conv(up, t, make_epilogue(bias=up.bias, res=h, affine=tabs["relu4_1"]))
            for block, lvl in (("slice4", "relu3_1"), ("slice3", "relu2_1"), ("slice2", "relu1_1")):
                bw = self.w[block]
                s = self._conv(bw["short"], h, make_epilogue())
                r1, part = self._conv_stats(bw["conv1"], h, make_epilogue(bias=bw["conv1"].bias, act=2), False)
                y = self._pointwise(r1, make_epilogue(norm1=self._finalize(part, 3, 1e-8)))
                del r1
                r2, part = self._conv_stats(bw["conv2"], y, make_epilogue(bias=bw["conv2"].bias, act=2), False)
                r, part = self._pointwise_stats(r2, make_epilogue(norm1=self._finalize(part, 3, 1e-8), res=s, res_shift=1), False)
                del r2
                h = self._pointwise(r, make_epilogue(norm1=self._finalize(part, 3, 1e-8), affine=tabs[lvl]))    # Decoder.AdaIN :311-319
                del r
            self._head(h, out[i:i + 1])
        return out

    @torch.no_grad()
    def forward_frame_graphed(self, frame, kind=0, gray=True):
        """forward_frame() replayed from a CUDA graph captured once per input shape and style: the ~90 launches of a frame-mode
        frame (statistics, filter prediction, rrv_fold_filter rewriting the filters' weight blobs in place, convolutions) become
        one graph launch.  Returns the graph's static output (valid until the next call)."""
        if self.style is None:
            raise RuntimeError("forward() before generate_style_features()")
        key = ("frame-graph", kind, bool(gray), tuple(frame.shape), frame.dtype)
        plan = self._plans.get(key)
        if plan is None:
            g_in = torch.empty_like(frame, memory_format=torch.contiguous_format)
            g_in.copy_(frame)
            n0 = self.lib.rrv_launch_count()
            g_out = self.forward_frame(g_in, kind, gray)                # eager warm-up: function attributes, caches, allocator pools
            launches = self.lib.rrv_launch_count() - n0
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.forward_frame(g_in, kind, gray, out=g_out)
            plan = (graph, g_in, g_out, launches)
            self._plans[key] = plan
            while len(self._plans) > self.MAX_PLANS:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        graph, g_in, g_out, launches = plan
        g_in.copy_(frame, non_blocking=True)
        graph.replay()
        self.graph_launches += launches
        return g_out

    def _folded(self, f, wf1, wf2):
        """The two convolutions of a KernelFilter with its predicted 32x32 matrices folded in (see _fold_filter), without
        touching the per-clip cache.  Tensor-core path: one kernel rewriting this filter's preallocated blobs."""
        fw = self.w[f]
        if self.impl_name != "ffma":
            cache = self.__dict__.setdefault("_frame_folds", {})
            if f not in cache or cache[f]._src[0].data_ptr() != fw["down_w"].data_ptr():
                cache[f] = FoldedFilter(fw, self.device)
            return cache[f].fold(wf1, wf2)
        dw = torch.matmul(wf1, fw["down_w"].reshape(32, -1)).reshape(32, 512, 3, 3)
        db = torch.mv(wf1, fw["down_b"])
        uw = torch.einsum("ojyx,ji->oiyx", fw["up_w"], wf2).contiguous()
        return ConvW(dw, db, cout_pad=INNER_PAD), ConvW(uw, fw["up_b"], cin_pad=INNER_PAD)

    @torch.no_grad()
    def vgg_features(self, x, top="Vgg19", kind=0):
        """Vgg19.forward of train/style_networks.py:284-314 (and EncoderStyle's four taps): relu1_1, relu2_1, relu3_1,
        relu4_1 as fp32 NCHW views of NHWC tensors, for a batch of normalised RGB images."""
        if top not in self.w:
            raise RuntimeError(f"no weights loaded under '{top}.*'")
        if kind == 0:
            N, _, H, W = x.shape
        else:
            N, H, W, _ = x.shape
        taps = self._vgg(top, x.contiguous(), kind, False, N, H, W, "style")
        return tuple(taps[i].permute(0, 3, 1, 2) for i in (0, 5, 10, 19))

    # ------------------------------------------------------------------ backward of the frozen Vgg19 loss network
    def _dgrad_weights(self, top):
        """Data-gradient weights of the nine VGG convolutions: W'[ci][co][ky][kx] = W[co][ci][2-ky][2-kx] (channel axes swapped,
        taps rotated by 180 degrees), packed once; the data gradient of a stride-1 'same' convolution is then rrv_conv2d itself."""
        key = top + "/dgrad"
        if key not in self.w:
            convs = self.w[top]
            flip = lambda w: w.transpose(0, 1).flip(2, 3).contiguous()
            self.w[key] = [ConvW(flip(convs[0][0]), None)] + [ConvW(flip(cw._keep[:cw.Cout, :cw.Cin]), None) for cw in convs[1:]]
        return self.w[key]

    @torch.no_grad()
    def vgg_features_train(self, x, top="Vgg19"):
        """Vgg19.forward (train/style_networks.py:284-314) keeping what its backward needs: every ReLU output in fp32 NHWC.
        Returns ((relu1_1, relu2_1, relu3_1, relu4_1) as NCHW views, saved)."""
        if top not in self.w:
            raise RuntimeError(f"no weights loaded under '{top}.*'")
        N, _, H, W = x.shape
        convs = self.w[top]
        w0, b0 = convs[0]
        cur = Planes(N, H, W, 64, self.x3, self.device)
        y = torch.empty((N, H, W, 64), dtype=torch.float32, device=self.device)
        L.check(self.lib.rrv_first_layer(x.contiguous().data_ptr(), 0, 0, N, H, W, w0.data_ptr(), b0.data_ptr(), L.ptr(cur.hi), L.ptr(cur.lo),
                                         y.data_ptr(), L.stream()), "rrv_first_layer")
        saved = [y]
        ident = make_epilogue()
        for (idx, _, _), cw in zip(VGG_CONVS[1:], convs[1:]):
            y = self._conv(cw, cur, make_epilogue(bias=cw.bias, act=1), L.OUT_F32_NHWC)
            saved.append(y)
            if idx != 19:
                cur = self._pointwise(y, ident)
                if idx in VGG_POOL_AFTER:
                    cur = self._pool(cur)
        feats = tuple(saved[i].permute(0, 3, 1, 2) for i in (0, 2, 4, 8))
        return feats, saved

    @torch.no_grad()
    def vgg_backward(self, saved, grads, top="Vgg19"):
        """d(loss)/d(input image) of the frozen Vgg19 given d(loss)/d(relu1_1, relu2_1, relu3_1, relu4_1) (NCHW, any may be None):
        per layer, ReLU backward fused with the split into operand planes (rrv_relu_backward), the data-gradient convolution on
        the tensor cores, and rrv_maxpool2x2_backward where the forward pooled.  Returns fp32 NCHW [N, 3, H, W]."""
        dw = self._dgrad_weights(top)
        taps = {0: 0, 2: 1, 4: 2, 8: 3}                     # layer position -> feature index
        pooled_after = {1, 3, 7}                            # conv1_2, conv2_2, conv3_4
        g = None
        for i in range(8, -1, -1):
            y = saved[i]
            N, H, W, Cc = y.shape
            if i in pooled_after and g is not None:
                gy = torch.empty_like(y)
                L.check(self.lib.rrv_maxpool2x2_backward(g.data_ptr(), y.data_ptr(), N, H, W, Cc, gy.data_ptr(), L.stream()),
                        "rrv_maxpool2x2_backward")
                g = gy
            gt = grads[taps[i]] if i in taps else None
            if gt is not None:
                gt = gt.permute(0, 2, 3, 1).contiguous().float()
                if tuple(gt.shape) != tuple(y.shape):
                    raise ValueError(f"gradient of feature {taps[i]} is {tuple(gt.shape)}, expected NCHW of {tuple(y.shape)}")
            if g is None and gt is None:
                continue
            a, b = (g, gt) if g is not None else (gt, None)
            gp = Planes(N, H, W, Cc, self.x3, self.device)
            L.check(self.lib.rrv_relu_backward(a.data_ptr(), L.ptr(b), y.data_ptr(), y.numel(), L.ptr(gp.hi), L.ptr(gp.lo), L.stream()),
                    "rrv_relu_backward")
            if i == 0:
                return self._conv(dw[0], gp, make_epilogue(), L.OUT_F32_NCHW, out_C=3)
            g = self._conv(dw[i], gp, make_epilogue(), L.OUT_F32_NHWC)
        N, H, W, _ = saved[0].shape
        return torch.zeros((N, 3, H, W), dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def feature_mean_std(self, feat_nchw_view):
        """calc_mean_std (train/style_networks.py:95-103) per sample: ([N,C] mean, [N,C] sqrt(unbiased var + 1e-5))."""
        x = feat_nchw_view.permute(0, 2, 3, 1)
        assert x.is_contiguous(), "expects the NCHW view of an NHWC tensor as returned by vgg_features"
        N, Hh, Ww, Cc = x.shape
        means, stds = [], []
        for i in range(N):
            part = torch.empty((5, Cc), dtype=torch.float64, device=self.device)
            L.check(self.lib.rrv_channel_stats(x[i].data_ptr(), Hh * Ww, Cc, part.data_ptr(), L.stream()), "rrv_channel_stats")
            t = self._finalize(part, 1, 1e-5)
            stds.append(t[0])
            means.append(t[1])
        return torch.stack(means), torch.stack(stds)

    # ------------------------------------------------------------------ state export (tests, dist)
    def export_clip_state(self):
        return dict(stats={k: v.clone() for k, v in self.stats.items()},
                    filters={k: (a.clone(), b.clone()) for k, (a, b) in self.filters.items()})

    def import_clip_state(self, state):
        self.stats = {k: v.to(self.device).contiguous() for k, v in state["stats"].items()}
        for f, (a, b) in state["filters"].items():
            self._fold_filter(f, a.to(self.device), b.to(self.device))
        self._plans = OrderedDict()
