"""Drop-in ``TransformerNet`` of the multi-style interpolation demo (reference:
``Multi-style Interpolation/style_network.py:464-507``).

The reference keeps, for every style, its own 11 saved-statistic tables, 6 dynamic filters and 4 (mean, std)
pairs, and blends them with ``style_weight`` inside every ``forward`` (InstanceNorm.forward :35-53,
FilterPredictor.forward :135-139, Decoder.AdaIN :348-360).  Only those small tables depend on the weights, so
here the blend is a few table additions on the device and the frame then runs through the same kernels as the
single-style path (csrc/conv_tc.cu).  Same 107-key ``state_dict`` and the same call sequence:

    net.generate_style_features(style, style_id)            # once per style
    net.add_patch(net.generate_content_features(patch))     # sampled frames of the clip
    net.compute_norm()                                      # per-style statistics (quirk Q1 applies per style)
    out = net(net.generate_content_features(frame), style_weight=[0.3, 0.7])
"""
from __future__ import annotations

import torch

from .style_network_global import TransformerNet as _GlobalNet


class TransformerNet(_GlobalNet):
    def __init__(self, style_num=1, precision="x3", impl="auto"):
        super().__init__(precision=precision, impl=impl)
        self.style_num = int(style_num)
        self.F_style = [None] * self.style_num
        self.F_patches = []
        self._styles = [None] * self.style_num          # engine-side style features (tables, normalised map)
        self._states = None                             # per style: {"stats", "filters"} after compute_norm()
        self._blended_for = None

    # ---- reference API (:475-497) ----
    def generate_style_features(self, style, style_id):
        eng = self._eng()
        eng.generate_style_features(style)
        self._styles[style_id], self.F_style[style_id] = eng.style, eng.F_style
        self._blended_for = None

    def generate_content_features(self, content):
        """Encoder(RGB2Gray(content)): [N,512,h/8,w/8] fp32 (an NCHW view of the NHWC tensor the kernels produce)."""
        return self._eng().encode(content)

    def add_patch(self, F_patch):
        self.F_patches.append(F_patch)

    def compute_norm(self):
        """Decoder.compute_norm (:415-430): the single-style pre-pass once per style on the same sampled features."""
        if any(s is None for s in self._styles):
            raise RuntimeError("compute_norm() before generate_style_features() of every style")
        if not self.F_patches:
            raise RuntimeError("compute_norm() without add_patch()")
        eng = self._eng()
        samples = [p.permute(0, 2, 3, 1).contiguous() for p in self.F_patches]
        self._states = []
        for sid in range(self.style_num):
            eng.clean()
            eng.style = self._styles[sid]
            eng.samples = list(samples)
            eng.compute(keep_samples=True)
            self._states.append(eng.export_clip_state())
        eng.samples = []
        self.F_patches = []
        self._blended_for = None

    def clean(self):
        """Decoder.clean (:399-413): drops the cached statistics and filters (the collected patches stay, like in the reference)."""
        self._states = None
        self._blended_for = None
        if self._engine is not None:
            self._engine.clean()

    def add(self, patch):                      # the single-style names do not exist on the reference's multi-style class
        raise AttributeError("'TransformerNet' (multi-style) object has no attribute 'add' (use add_patch)")

    def compute(self):
        raise AttributeError("'TransformerNet' (multi-style) object has no attribute 'compute' (use compute_norm)")

    def _blend(self, style_weight):
        w = [float(x) for x in style_weight]
        if len(w) != self.style_num:
            raise ValueError(f"style_weight needs {self.style_num} entries, got {len(w)}")
        key = tuple(w)
        if self._blended_for == key:
            return
        if self._states is None:
            raise RuntimeError("forward() before compute_norm(): the per-style statistics are not available")
        eng = self._eng()
        mix = lambda ts: sum(wi * t for wi, t in zip(w, ts))
        state = dict(stats={k: mix([st["stats"][k] for st in self._states]) for k in self._states[0]["stats"]},
                     filters={k: (mix([st["filters"][k][0] for st in self._states]), mix([st["filters"][k][1] for st in self._states]))
                              for k in self._states[0]["filters"]})
        # AdaIN tables {std, mean} per level (:348-360); the normalised style map is only needed by compute_norm
        tabs = {lvl: mix([s["tabs"][lvl] for s in self._styles]).contiguous() for lvl in self._styles[0]["tabs"]}
        eng.style = dict(tabs=tabs, map=None, nstyle=None)
        eng.import_clip_state(state)
        self._blended_for = key

    def forward(self, F_content, style_weight=(1.0,)):
        """Decoder(F_content, style_weight) (:478-484): F_content from generate_content_features."""
        self._blend(style_weight)
        return self._eng().decode_features(F_content)
