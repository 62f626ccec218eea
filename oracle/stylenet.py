"""CPU oracle for the ReReVST per-frame stylization path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
(``rerevst-code_b200``) never does.

It is a functional restatement, on CPU tensors, of what the reference computes,
written against a plain 107-key ``state_dict``.  All arithmetic goes through the
same third-party library the reference uses (PyTorch CPU ops: ``F.conv2d``,
``F.max_pool2d``, ``F.interpolate``, ``torch.mean/max/min/rsqrt``), so on
identical weights and inputs it reproduces the reference bit for bit; this is
pinned by ``tests/test_oracle.py::test_oracle_equals_live_reference`` (runs where
``/root/reference`` exists) and by the committed fixtures under ``tests/golden/`` that
``oracle/make_golden.py`` generated from the unmodified reference modules.

Reference files restated here (paths relative to the reference repo):
  test/style_network_global.py   global ("Sequence-Level Global Feature Sharing") mode
  test/style_network_frame.py    frame mode (per-frame statistics)
  test/framework.py              uint8 BGR <-> normalised tensor conversions
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

MeanStd = namedtuple("MeanStd", ["mean", "std"])
StyleFeatures = namedtuple("StyleFeatures", ["map", "relu1_1", "relu2_1", "relu3_1", "relu4_1"])

_VGG_IDX = (0, 2, 5, 7, 10, 12, 14, 16, 19)
_POOL_AFTER = (2, 7, 16)
_STYLE_SLICE = {0: "slice1", 2: "slice2", 5: "slice2", 7: "slice3", 10: "slice3",
                12: "slice4", 14: "slice4", 16: "slice4", 19: "slice4"}
_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------
# test/framework.py:26-49 -- image <-> tensor

def numpy2tensor(img_bgr_u8: np.ndarray) -> torch.Tensor:
    """framework.py:26-28: BGR HWC uint8 -> RGB CHW float32 (no scaling)."""
    rgb = np.ascontiguousarray(img_bgr_u8[:, :, ::-1])
    return torch.from_numpy(rgb.transpose((2, 0, 1)).copy()).float()


def transform_image(img: torch.Tensor) -> torch.Tensor:
    """framework.py:30-35: /255, ImageNet normalise, add batch dim."""
    mean = img.new_tensor(_IMAGENET_MEAN).view(-1, 1, 1)
    std = img.new_tensor(_IMAGENET_STD).view(-1, 1, 1)
    img = img.div(255.0)
    img = (img - mean) / std
    return img.unsqueeze(0)


def transform_back_image(img: torch.Tensor) -> torch.Tensor:
    """framework.py:44-49: de-normalise, clamp to [0,1], first batch item, x255."""
    mean = img.new_tensor(_IMAGENET_MEAN).view(-1, 1, 1)
    std = img.new_tensor(_IMAGENET_STD).view(-1, 1, 1)
    img = img * std + mean
    return img.clamp(0, 1)[0, :, :, :] * 255


def tensor2numpy(img: torch.Tensor) -> np.ndarray:
    """framework.py:39-42: CHW RGB float -> HWC BGR float32."""
    arr = img.detach().cpu().numpy().transpose((1, 2, 0))
    return np.ascontiguousarray(arr[:, :, ::-1])


# --------------------------------------------------------------------------------------
# test/style_network_global.py:487-497 -- RGB2Gray (BGR weights applied to an RGB tensor, kept)

def rgb2gray(image: torch.Tensor) -> torch.Tensor:
    mean = image.new_tensor(_IMAGENET_MEAN).view(-1, 1, 1)
    std = image.new_tensor(_IMAGENET_STD).view(-1, 1, 1)
    image = image * std + mean
    gray = image[:, 2:3] * 0.299 + image[:, 1:2] * 0.587 + image[:, 0:1] * 0.114
    gray = gray.expand(image.size())
    return (gray - mean) / std


# --------------------------------------------------------------------------------------
# VGG-19 features[0:21]  (style_network_global.py:271-281, 284-331, 238-268)

def _vgg_stack(x, sd, names, taps=()):
    """Nine 3x3 conv+ReLU with 2x2 max-pools after conv1_2, conv2_2, conv3_4.
    ``taps``: feature indices after whose ReLU the activation is also returned."""
    outs = {}
    for idx, (wk, bk) in zip(_VGG_IDX, names):
        x = F.relu(F.conv2d(x, sd[wk], sd[bk], padding=1))
        if idx in taps:
            outs[idx] = x
        if idx in _POOL_AFTER:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
    return x, outs


def _enc_names(top):
    if top == "Encoder":
        return [(f"Encoder.slice.{i}.weight", f"Encoder.slice.{i}.bias") for i in _VGG_IDX]
    return [(f"{top}.{_STYLE_SLICE[i]}.{i}.weight", f"{top}.{_STYLE_SLICE[i]}.{i}.bias") for i in _VGG_IDX]


def encoder(x, sd):
    """Encoder.forward, style_network_global.py:280-281."""
    return _vgg_stack(x, sd, _enc_names("Encoder"))[0]


def vgg19_features(x, sd, top="Vgg19"):
    """Vgg19.forward, style_network_global.py:258-268 (relu1_1, 2_1, 3_1, 4_1 maps)."""
    _, o = _vgg_stack(x, sd, _enc_names(top), taps=(0, 5, 10, 19))
    return o[0], o[5], o[10], o[19]


def cal_mean_std(feat, eps=1e-5):
    """EncoderStyle.cal_mean_std, style_network_global.py:304-315 (unbiased var)."""
    n, c = feat.shape[:2]
    var = feat.reshape(n, c, -1).var(dim=2) + eps
    std = var.sqrt().view(n, c, 1, 1)
    mean = feat.reshape(n, c, -1).mean(dim=2).view(n, c, 1, 1)
    return MeanStd(mean, std)


def encoder_style(style, sd):
    """EncoderStyle.forward, style_network_global.py:317-331."""
    f1, f2, f3, f4 = vgg19_features(style, sd, top="EncoderStyle")
    return StyleFeatures(f4, cal_mean_std(f1), cal_mean_std(f2), cal_mean_std(f3), cal_mean_std(f4))


# --------------------------------------------------------------------------------------
# Saved-statistic InstanceNorm  (style_network_global.py:27-84)

SavedStat = namedtuple("SavedStat", ["mean", "rstd", "lo", "hi"])


def in_compute(x, eps=1e-8):
    """InstanceNorm.compute, :59-77.  Statistics over batch AND space, biased two-pass
    variance, min/max of the normalised tensor.  Returns (stats, unclamped x_hat)."""
    mean = torch.mean(x, (0, 2, 3), True)
    x = x - mean
    rstd = torch.rsqrt(torch.mean(torch.mul(x, x), (0, 2, 3), True) + eps)
    x = x * rstd
    hi = x.amax(dim=(0, 2, 3), keepdim=True)
    lo = x.amin(dim=(0, 2, 3), keepdim=True)
    return SavedStat(mean, rstd, lo, hi), x


def in_forward(x, st: SavedStat):
    """InstanceNorm.forward, :43-57."""
    x = x - st.mean
    x = x * st.rstd
    x = torch.max(st.lo, x)
    x = torch.min(st.hi, x)
    return x


def in_frame(x, eps=1e-8):
    """Frame-mode InstanceNorm.forward, style_network_frame.py:39-43 (per sample, no clamp)."""
    x = x - torch.mean(x, (2, 3), True)
    r = torch.rsqrt(torch.mean(torch.mul(x, x), (2, 3), True) + eps)
    return x * r


# --------------------------------------------------------------------------------------
# KernelFilter / FilterPredictor  (style_network_global.py:142-230)

def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def _predict_filter(sd, pfx, content, style, batch_mean):
    """FilterPredictor.compute (:161-172) when batch_mean, .forward (:150-159) otherwise.
    Returns [B,32,32] with [b, out_channel, in_channel] (the permute at :205 makes the
    first view dim the conv output channel)."""
    w, b = sd[pfx + ".down_sample.0.weight"], sd[pfx + ".down_sample.0.bias"]
    c = F.conv2d(content, w, b, padding=1)
    c = torch.mean(c.reshape(c.size(0), c.size(1), -1), dim=2)
    if batch_mean:
        c = torch.mean(c, dim=0).unsqueeze(0)
    s = F.conv2d(style, w, b, padding=1)
    s = torch.mean(s.reshape(s.size(0), s.size(1), -1), dim=2)
    f = F.linear(torch.cat([c, s], 1), sd[pfx + ".FC.weight"], sd[pfx + ".FC.bias"])
    return f.view(-1, 32, 32)


def _apply_filter(x, filt):
    """KernelFilter.apply_filter, :194-208.  ``zip`` over the chunked input and the chunked
    filter stops at the shorter one: with a batch-1 filter only sample 0 survives (quirk Q1)."""
    n = min(x.shape[0], filt.shape[0])
    outs = [F.conv2d(x[i:i + 1], filt[i].unsqueeze(-1).unsqueeze(-1)) for i in range(n)]
    return torch.cat(outs, 0)


def kernel_filter(sd, pfx, content, wf1, wf2):
    """KernelFilter.forward with cached filters, :210-217."""
    t = F.conv2d(content, sd[pfx + ".down_sample.0.weight"], sd[pfx + ".down_sample.0.bias"], padding=1)
    t = _apply_filter(t, wf1)
    t = _lrelu(t)
    t = _apply_filter(t, wf2)
    return content + F.conv2d(t, sd[pfx + ".upsample.0.weight"], sd[pfx + ".upsample.0.bias"], padding=1)


def kernel_filter_compute(sd, pfx, content, style):
    """KernelFilter.compute, :223-230.  Returns (out, wf1, wf2)."""
    t = F.conv2d(content, sd[pfx + ".down_sample.0.weight"], sd[pfx + ".down_sample.0.bias"], padding=1)
    wf1 = _predict_filter(sd, pfx + ".F1", content, style, batch_mean=True)
    t = _apply_filter(t, wf1)
    t = _lrelu(t)
    wf2 = _predict_filter(sd, pfx + ".F2", content, style, batch_mean=True)
    t = _apply_filter(t, wf2)
    out = content + F.conv2d(t, sd[pfx + ".upsample.0.weight"], sd[pfx + ".upsample.0.bias"], padding=1)
    return out, wf1, wf2


# --------------------------------------------------------------------------------------
# ResidualBlock  (style_network_global.py:100-139)

def _res_convs(sd, pfx, x):
    x = F.interpolate(x, mode="nearest", scale_factor=2)
    xs = F.conv2d(x, sd[pfx + ".conv_shortcut.weight"])
    return x, xs


def residual_block(sd, pfx, x, st1, st2):
    x, xs = _res_convs(sd, pfx, x)
    x = in_forward(_lrelu(F.conv2d(x, sd[pfx + ".conv1.weight"], sd[pfx + ".conv1.bias"], padding=1)), st1)
    x = in_forward(_lrelu(F.conv2d(x, sd[pfx + ".conv2.weight"], sd[pfx + ".conv2.bias"], padding=1)), st2)
    return xs + x


def residual_block_compute(sd, pfx, x):
    x, xs = _res_convs(sd, pfx, x)
    st1, x = in_compute(_lrelu(F.conv2d(x, sd[pfx + ".conv1.weight"], sd[pfx + ".conv1.bias"], padding=1)))
    st2, x = in_compute(_lrelu(F.conv2d(x, sd[pfx + ".conv2.weight"], sd[pfx + ".conv2.bias"], padding=1)))
    return xs + x, st1, st2


def residual_block_frame(sd, pfx, x):
    """Frame-mode ResidualBlock.forward, style_network_frame.py:179-192."""
    x, xs = _res_convs(sd, pfx, x)
    x = in_frame(_lrelu(F.conv2d(x, sd[pfx + ".conv1.weight"], sd[pfx + ".conv1.bias"], padding=1)))
    x = in_frame(_lrelu(F.conv2d(x, sd[pfx + ".conv2.weight"], sd[pfx + ".conv2.bias"], padding=1)))
    return xs + x


# --------------------------------------------------------------------------------------
# Decoder / TransformerNet, global mode

class ClipState:
    """What the reference caches on its modules between ``compute()`` and ``forward``:
    11 saved-stat tables and 6 dynamic filters.  Order of ``stats``:
      norm0, norm1, slice4.norm1, slice4.norm2, norm2, slice3.norm1, slice3.norm2,
      norm3, slice2.norm1, slice2.norm2, norm4
    (Decoder.norm[i] at style_network_global.py:347-351, block norms at :107-108)."""

    NAMES = ("norm0", "norm1", "slice4.norm1", "slice4.norm2", "norm2", "slice3.norm1",
             "slice3.norm2", "norm3", "slice2.norm1", "slice2.norm2", "norm4")

    def __init__(self):
        self.stats = {}
        self.filters = {}


def decoder_compute(sd, x, fs: StyleFeatures) -> ClipState:
    """Decoder.compute, style_network_global.py:425-439 (with AdaIN_filter_compute :392-402,
    AdaIN_compute :383-390).  ``x`` is the concatenated encoder output of the sampled frames."""
    cs = ClipState()
    st, h = in_compute(x)
    cs.stats["norm0"] = st
    nstyle = (fs.map - fs.relu4_1.mean) / fs.relu4_1.std
    for f in ("Filter1", "Filter2", "Filter3"):
        h, wf1, wf2 = kernel_filter_compute(sd, "Decoder." + f, h, nstyle)
        cs.filters[f] = (wf1, wf2)
    levels = (("norm1", fs.relu4_1, "slice4"), ("norm2", fs.relu3_1, "slice3"),
              ("norm3", fs.relu2_1, "slice2"), ("norm4", fs.relu1_1, None))
    for nname, sf, block in levels:
        st, h = in_compute(h)
        cs.stats[nname] = st
        h = h * sf.std + sf.mean
        if block is not None:
            h, s1, s2 = residual_block_compute(sd, "Decoder." + block, h)
            cs.stats[block + ".norm1"] = s1
            cs.stats[block + ".norm2"] = s2
    return cs


def decoder_forward(sd, x, fs: StyleFeatures, cs: ClipState, taps=None):
    """Decoder.forward, style_network_global.py:441-451."""
    h = in_forward(x, cs.stats["norm0"])
    for f in ("Filter1", "Filter2", "Filter3"):
        h = kernel_filter(sd, "Decoder." + f, h, *cs.filters[f])
        if taps is not None:
            taps[f] = h
    levels = (("norm1", fs.relu4_1, "slice4"), ("norm2", fs.relu3_1, "slice3"),
              ("norm3", fs.relu2_1, "slice2"), ("norm4", fs.relu1_1, None))
    for nname, sf, block in levels:
        h = in_forward(h, cs.stats[nname]) * sf.std + sf.mean
        if block is not None:
            h = residual_block(sd, "Decoder." + block, h, cs.stats[block + ".norm1"], cs.stats[block + ".norm2"])
            if taps is not None:
                taps[block] = h
    return F.conv2d(h, sd["Decoder.slice1.weight"], sd["Decoder.slice1.bias"], padding=1)


class GlobalOracle:
    """Same call sequence as the reference ``TransformerNet`` in global mode
    (style_network_global.py:454-501): generate_style_features, clean, add, compute, forward."""

    def __init__(self, state_dict, device="cpu"):
        """device="cpu" is the oracle.  Any other device runs the SAME functional code through PyTorch's eager kernels there
        (cuDNN on CUDA): bench.py's ``gpu_baseline`` -- what the reference itself does on a GPU (test/framework.py:61-65)."""
        self.device = torch.device(device)
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        self.F_style = None
        self.clip = None
        self.F_patches = None

    def to_device_state(self, clip: "ClipState", fs: "StyleFeatures"):
        """Adopt per-clip tables / style statistics computed elsewhere (moved to this oracle's device)."""
        mv = lambda t: None if t is None else t.to(self.device)
        c = ClipState()
        c.stats = {k: SavedStat(*[mv(t) for t in v]) for k, v in clip.stats.items()}
        c.filters = {k: tuple(mv(t) for t in v) for k, v in clip.filters.items()}
        self.clip = c
        self.F_style = StyleFeatures(mv(fs.map), *[MeanStd(mv(m.mean), mv(m.std)) for m in fs[1:]])

    @torch.no_grad()
    def generate_style_features(self, style):
        self.F_style = encoder_style(style, self.sd)

    def clean(self):
        self.F_patches = []
        self.clip = None

    @torch.no_grad()
    def add(self, patch):
        self.F_patches.append(encoder(rgb2gray(patch), self.sd))

    @torch.no_grad()
    def compute(self):
        self.clip = decoder_compute(self.sd, torch.cat(self.F_patches, dim=0), self.F_style)

    @torch.no_grad()
    def forward(self, frame, taps=None):
        fc = encoder(rgb2gray(frame), self.sd)
        if taps is not None:
            taps["F_content"] = fc
        return decoder_forward(self.sd, fc, self.F_style, self.clip, taps)

    __call__ = forward

    # framework.Stylization.transfer, test/framework.py:106-118
    @torch.no_grad()
    def transfer(self, frame_bgr_u8):
        x = transform_image(numpy2tensor(frame_bgr_u8))
        return tensor2numpy(transform_back_image(self.forward(x)))


# --------------------------------------------------------------------------------------
# Multi-style interpolation  ("Multi-style Interpolation/style_network.py")

def blend_states(states, styles, weights):
    """What the multi-style modules do inside forward (InstanceNorm.forward :35-53, FilterPredictor.forward :135-139,
    Decoder.AdaIN :348-360): every cached table is the style_weight-ed sum of the per-style tables."""
    cs = ClipState()
    for k in states[0].stats:
        cs.stats[k] = SavedStat(*[sum(w * getattr(st.stats[k], f) for st, w in zip(states, weights)) for f in SavedStat._fields])
    for k in states[0].filters:
        cs.filters[k] = tuple(sum(w * st.filters[k][j] for st, w in zip(states, weights)) for j in range(2))
    blend = lambda lvl: MeanStd(sum(w * getattr(fs, lvl).mean for fs, w in zip(styles, weights)),
                                sum(w * getattr(fs, lvl).std for fs, w in zip(styles, weights)))
    fs = StyleFeatures(None, blend("relu1_1"), blend("relu2_1"), blend("relu3_1"), blend("relu4_1"))
    return cs, fs


class MultiStyleOracle:
    """Call sequence of the multi-style ``TransformerNet`` (:464-497): generate_style_features(style, id),
    generate_content_features, add_patch, compute_norm, forward(F_content, style_weight)."""

    def __init__(self, state_dict, style_num):
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self.F_style = [None] * style_num
        self.F_patches = []
        self.states = None

    @torch.no_grad()
    def generate_style_features(self, style, style_id):
        self.F_style[style_id] = encoder_style(style, self.sd)

    @torch.no_grad()
    def generate_content_features(self, content):
        return encoder(rgb2gray(content), self.sd)

    def add_patch(self, f_patch):
        self.F_patches.append(f_patch)

    @torch.no_grad()
    def compute_norm(self):
        x = torch.cat(self.F_patches, dim=0)
        self.states = [decoder_compute(self.sd, x, fs) for fs in self.F_style]       # Decoder.compute_norm :415-430, per style
        self.F_patches = []

    @torch.no_grad()
    def forward(self, f_content, style_weight=(1.0,)):
        cs, fs = blend_states(self.states, self.F_style, style_weight)
        return decoder_forward(self.sd, f_content, fs, cs)


# --------------------------------------------------------------------------------------
# Frame mode  (style_network_frame.py:294-394)

@torch.no_grad()
def frame_mode_forward(sd, frame, fs: StyleFeatures, gray=True):
    """TransformerNet.forward in frame mode (style_network_frame.py:392-394, Decoder.forward
    :341-358).  With ``gray=False`` it is ``validation`` of train/style_networks.py:556-559."""
    x = encoder(rgb2gray(frame) if gray else frame, sd)
    mean, std = fs.relu4_1.mean, fs.relu4_1.std
    h = in_frame(x)
    nstyle = (fs.map - mean) / std
    for f in ("Filter1", "Filter2", "Filter3"):
        pfx = "Decoder." + f
        t = F.conv2d(h, sd[pfx + ".down_sample.0.weight"], sd[pfx + ".down_sample.0.bias"], padding=1)
        t = _apply_filter(t, _predict_filter(sd, pfx + ".F1", h, nstyle, batch_mean=False))
        t = _lrelu(t)
        t = _apply_filter(t, _predict_filter(sd, pfx + ".F2", h, nstyle, batch_mean=False))
        h = h + F.conv2d(t, sd[pfx + ".upsample.0.weight"], sd[pfx + ".upsample.0.bias"], padding=1)
    h = h * std + mean
    h = residual_block_frame(sd, "Decoder.slice4", h)
    h = in_frame(h) * fs.relu3_1.std + fs.relu3_1.mean
    h = residual_block_frame(sd, "Decoder.slice3", h)
    h = in_frame(h) * fs.relu2_1.std + fs.relu2_1.mean
    h = residual_block_frame(sd, "Decoder.slice2", h)
    h = in_frame(h) * fs.relu1_1.std + fs.relu1_1.mean
    return F.conv2d(h, sd["Decoder.slice1.weight"], sd["Decoder.slice1.bias"], padding=1)


# --------------------------------------------------------------------------------------
# Work model: algorithmic FLOPs of the 31 convolutions per frame

def conv_flops_per_frame(h: int, w: int) -> float:
    """2*Cin*Cout*k*k*Hout*Wout over the 9 encoder, 12 KernelFilter... see BASELINE.md:3.
    (The 6 FilterPredictor convs are pre-pass only and not counted.)"""
    total = 0.0
    hh, ww = h, w
    for idx, cin, cout in ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256),
                           (12, 256, 256), (14, 256, 256), (16, 256, 256), (19, 256, 512)):
        total += 2.0 * cin * cout * 9 * hh * ww
        if idx in _POOL_AFTER:
            hh, ww = hh // 2, ww // 2
    total += 3 * (2.0 * 512 * 32 * 9 + 2 * 2.0 * 32 * 32 + 2.0 * 32 * 512 * 9) * hh * ww
    for cin, cout in ((512, 256), (256, 128), (128, 64)):
        hh, ww = hh * 2, ww * 2
        total += (2.0 * cin * cout * 9 + 2.0 * cout * cout * 9 + 2.0 * cin * cout) * hh * ww
    total += 2.0 * 64 * 3 * 9 * hh * ww
    return total
