"""Drop-in ``Stylization`` facade (reference: ``test/framework.py:56-118``).

Same constructor and the same six methods; frames cross the boundary as uint8 BGR HWC numpy
arrays and come back as float32 BGR HWC in [0, 255].  The uint8 frame is uploaded as is (1/4 of
the reference's fp32 upload) and numpy2tensor/transform_image (:26-35) are fused into the first
convolution kernel; transform_back_image/tensor2numpy (:39-49) are one kernel on the way out.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


class Stylization:
    def __init__(self, checkpoint, cuda=True, use_Global=True, precision="x3", impl="auto"):
        if not cuda:
            raise RuntimeError("rerevst_b200.Stylization needs cuda=True (the reference's CPU path is the oracle, "
                               "not part of this package)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        if use_Global:
            from .style_network_global import TransformerNet
        else:
            from .style_network_frame import TransformerNet
        self.model = TransformerNet(precision=precision, impl=impl).to(self.device)
        sd = checkpoint if isinstance(checkpoint, dict) else torch.load(checkpoint, map_location="cpu")
        self.model.load_state_dict(sd)
        for p in self.model.parameters():
            p.requires_grad = False
        self._pin = {}

    def _upload(self, img):
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected a uint8 HxWx3 BGR image (cv2.imread layout)")
        key = img.shape
        if key not in self._pin:
            self._pin[key] = torch.empty((1,) + key, dtype=torch.uint8).pin_memory()
        self._pin[key][0].copy_(torch.from_numpy(img))
        return self._pin[key].to(self.device, non_blocking=True)

    # ===== Sequence-Level Global Feature Sharing =====
    def add(self, patch):
        eng = self.model._eng()
        if self.model.F_patches is None:
            raise AttributeError("call clean() before add()")
        eng.add(self._upload(patch), kind=1)
        self.model.F_patches.append(eng.samples[-1])

    def compute(self):
        self.model.compute()

    def clean(self):
        self.model.clean()

    # ===== Style Transfer =====
    def prepare_style(self, style):
        eng = self.model._eng()
        eng.generate_style_features(self._upload(style), kind=1)
        self.model.F_style = eng.F_style
        if not self.model.have_delete_vgg:
            del self.model.Vgg19
            self.model.have_delete_vgg = True

    def transfer(self, frame, crop=None):
        """frame: uint8 BGR HWC -> float32 BGR HWC in [0,255] (test/framework.py:106-118).
        crop=(y0, x0, h, w) additionally applies generate_real_video.py:167 on the device."""
        out = self.transfer_device(frame, crop)
        return out.cpu().numpy()[0]

    def transfer_device(self, frame, crop=None):
        eng = self.model._eng()
        y = eng.forward(self._upload(frame), kind=1)
        N, _, H, W = y.shape
        y0, x0, h, w = crop if crop is not None else (0, 0, H, W)
        out = torch.empty((N, h, w, 3), dtype=torch.float32, device=self.device)
        L.check(L.lib().rrv_postprocess_bgr(y.data_ptr(), N, H, W, y0, x0, h, w, out.data_ptr(), L.stream()),
                "rrv_postprocess_bgr")
        return out
