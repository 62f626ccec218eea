"""The 107-key ``state_dict`` contract of the reference ``TransformerNet`` and a
deterministic synthetic checkpoint.

Reference: ``test/style_network_global.py:454-463`` builds ``Decoder``,
``Encoder``, ``EncoderStyle`` and ``Vgg19``; ``test/framework.py:75`` loads the
checkpoint strictly, so every key below (including the unused ``Vgg19.*``
ones) has to be accepted.  Both checkpoints shipped with the reference are
empty placeholders, so tests and the benchmark use :func:`synthetic_state_dict`.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

# (index in vgg19().features, Cin, Cout) of the nine convolutions up to relu4_1
# (test/style_network_global.py:275-278 takes features[0:21]).
VGG_CONVS = (
    (0, 3, 64), (2, 64, 64),
    (5, 64, 128), (7, 128, 128),
    (10, 128, 256), (12, 256, 256), (14, 256, 256), (16, 256, 256),
    (19, 256, 512),
)
# 2x2 max-pools sit after these feature indices (vgg19 'M' entries 4, 9, 18).
VGG_POOL_AFTER = (2, 7, 16)

# EncoderStyle / Vgg19 split the same nine convs into four Sequentials
# (test/style_network_global.py:295-302).
_SLICE_OF = {0: "slice1", 2: "slice2", 5: "slice2", 7: "slice3", 10: "slice3",
             12: "slice4", 14: "slice4", 16: "slice4", 19: "slice4"}

RES_BLOCKS = (("slice4", 512, 256), ("slice3", 256, 128), ("slice2", 128, 64))
FILTERS = ("Filter1", "Filter2", "Filter3")
VGG_CH = 512
INNER_CH = 32


def key_shapes() -> "OrderedDict[str, tuple]":
    """All 107 parameter names with their shapes."""
    ks: "OrderedDict[str, tuple]" = OrderedDict()
    for name, cin, cout in RES_BLOCKS:
        ks[f"Decoder.{name}.conv1.weight"] = (cout, cin, 3, 3)
        ks[f"Decoder.{name}.conv1.bias"] = (cout,)
        ks[f"Decoder.{name}.conv2.weight"] = (cout, cout, 3, 3)
        ks[f"Decoder.{name}.conv2.bias"] = (cout,)
        ks[f"Decoder.{name}.conv_shortcut.weight"] = (cout, cin, 1, 1)
    ks["Decoder.slice1.weight"] = (3, 64, 3, 3)
    ks["Decoder.slice1.bias"] = (3,)
    for f in FILTERS:
        ks[f"Decoder.{f}.down_sample.0.weight"] = (INNER_CH, VGG_CH, 3, 3)
        ks[f"Decoder.{f}.down_sample.0.bias"] = (INNER_CH,)
        ks[f"Decoder.{f}.upsample.0.weight"] = (VGG_CH, INNER_CH, 3, 3)
        ks[f"Decoder.{f}.upsample.0.bias"] = (VGG_CH,)
        for p in ("F1", "F2"):
            ks[f"Decoder.{f}.{p}.down_sample.0.weight"] = (INNER_CH, VGG_CH, 3, 3)
            ks[f"Decoder.{f}.{p}.down_sample.0.bias"] = (INNER_CH,)
            ks[f"Decoder.{f}.{p}.FC.weight"] = (INNER_CH * INNER_CH, 2 * INNER_CH)
            ks[f"Decoder.{f}.{p}.FC.bias"] = (INNER_CH * INNER_CH,)
    for idx, cin, cout in VGG_CONVS:
        ks[f"Encoder.slice.{idx}.weight"] = (cout, cin, 3, 3)
        ks[f"Encoder.slice.{idx}.bias"] = (cout,)
    for top in ("EncoderStyle", "Vgg19"):
        for idx, cin, cout in VGG_CONVS:
            ks[f"{top}.{_SLICE_OF[idx]}.{idx}.weight"] = (cout, cin, 3, 3)
            ks[f"{top}.{_SLICE_OF[idx]}.{idx}.bias"] = (cout,)
    assert len(ks) == 107
    return ks


def vgg_keys(top: str):
    """[(weight_key, bias_key)] of the nine VGG convs under ``top``."""
    out = []
    for idx, _, _ in VGG_CONVS:
        if top == "Encoder":
            base = f"Encoder.slice.{idx}"
        else:
            base = f"{top}.{_SLICE_OF[idx]}.{idx}"
        out.append((base + ".weight", base + ".bias"))
    return out


def synthetic_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random fp32 weights for all 107 keys (CPU tensors).

    Scales are chosen so activations stay O(1) through the 31 convolutions:
    He-normal for the ReLU VGG stacks, 1/sqrt(fan_in)-uniform (PyTorch's
    Conv2d/Linear default bound) for the decoder, small non-zero biases so every
    bias path is exercised.  Each tensor has its own generator stream, so the
    values do not depend on construction order.
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, (k, shape) in enumerate(key_shapes().items()):
        g = torch.Generator().manual_seed(seed * 1000003 + 7919 * i + 17)
        if k.endswith("weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if k.startswith("Decoder"):
                bound = (3.0 / fan_in) ** 0.5
                if ".FC." in k:
                    bound = (1.0 / fan_in) ** 0.5
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            else:
                t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[k] = t.float().contiguous()
    return sd


def check_state_dict(sd) -> None:
    """Strict key/shape check, like ``load_state_dict`` at test/framework.py:75."""
    want = key_shapes()
    missing = [k for k in want if k not in sd]
    extra = [k for k in sd if k not in want]
    if missing or extra:
        raise RuntimeError(f"state_dict mismatch: missing={missing[:4]} unexpected={extra[:4]}")
    for k, shape in want.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
