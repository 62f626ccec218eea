"""Drop-in ``TransformerNet`` (global / "Sequence-Level Global Feature Sharing" mode).

Same constructor, same 107-key ``state_dict`` and same methods as the reference class
(``test/style_network_global.py:454-501``): ``generate_style_features``, ``add``, ``compute``,
``clean``, ``RGB2Gray``, ``forward``.  Every tensor op of the reference's ``Encoder``,
``EncoderStyle`` and ``Decoder`` runs in ``librerevst_b200.so`` (see engine.py); the module only
owns the parameters and the per-clip cache.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import StyleEngine
from .weights import key_shapes


class _Params(nn.Module):
    """A parameter-only module tree whose ``state_dict`` keys equal the reference's."""

    def __init__(self, prefix, shapes):
        super().__init__()
        children = {}
        for key, shape in shapes.items():
            if not key.startswith(prefix):
                continue
            rest = key[len(prefix):]
            head, _, tail = rest.partition(".")
            if tail == "":
                self.register_parameter(head, nn.Parameter(torch.zeros(shape), requires_grad=False))
            else:
                children.setdefault(head, None)
        for head in children:
            self.add_module(head, _Params(prefix + head + ".", shapes))


class TransformerNet(nn.Module):
    def __init__(self, precision="x3", impl="auto"):
        super().__init__()
        shapes = key_shapes()
        self.Decoder = _Params("Decoder.", shapes)
        self.Encoder = _Params("Encoder.", shapes)
        self.EncoderStyle = _Params("EncoderStyle.", shapes)
        self.Vgg19 = _Params("Vgg19.", shapes)      # accepted by load_state_dict, never used (:460, :467-469)
        self.have_delete_vgg = False
        self.precision, self.impl = precision, impl
        self._engine = None
        self._packed_for = None
        self.F_patches = None

    # ---- engine management ----
    def _eng(self) -> StyleEngine:
        p = self.Decoder.slice1.weight
        if p.device.type != "cuda":
            raise RuntimeError("rerevst_b200.TransformerNet computes on CUDA only: call .to('cuda') first "
                               "(there is no CPU fallback)")
        if self._engine is None or self._engine.device != p.device:
            self._engine = StyleEngine(p.device, self.precision, self.impl)
            self._packed_for = None
        sig = tuple((q.data_ptr(), q._version) for q in self.parameters())
        if self._packed_for != sig:
            sd = {k: v for k, v in self.state_dict().items()}
            if "Vgg19.slice1.0.weight" not in sd:          # deleted after the first style, like the reference
                sd = dict(sd)
            self._engine.load_weights(sd)
            self._packed_for = sig
        return self._engine

    # ---- reference API ----
    def generate_style_features(self, style):
        self._eng().generate_style_features(style)
        self.F_style = self._engine.F_style
        if not self.have_delete_vgg:
            del self.Vgg19
            self.have_delete_vgg = True

    def add(self, patch):
        if self.F_patches is None:
            raise AttributeError("'TransformerNet' object has no attribute 'F_patches' (call clean() before add(), "
                                 "test/style_network_global.py:484)")
        self._eng().add(patch)
        self.F_patches.append(self._engine.samples[-1])

    def compute(self):
        self._eng().compute()

    def clean(self):
        self.num = 0
        self.long_seq = False
        self.F_patches = []
        if self._engine is not None:
            self._engine.clean()

    def RGB2Gray(self, image):
        """Convenience copy of the reference method (:487-497); inside ``forward`` this conversion is
        fused into the first convolution kernel (csrc/first_layer.cu) and this method is not called."""
        mean = image.new_tensor([0.485, 0.456, 0.406]).view(-1, 1, 1)
        std = image.new_tensor([0.229, 0.224, 0.225]).view(-1, 1, 1)
        image = image * std + mean
        gray = image[:, 2:3] * 0.299 + image[:, 1:2] * 0.587 + image[:, 0:1] * 0.114
        return (gray.expand(image.size()) - mean) / std

    def forward(self, input_frame):
        return self._eng().forward(input_frame)

    # ---- extras ----
    def forward_u8(self, frame_bgr_u8):
        """uint8 NHWC BGR frames in, fp32 NCHW out: numpy2tensor + transform_image fused into the first kernel."""
        return self._eng().forward(frame_bgr_u8, kind=1)
