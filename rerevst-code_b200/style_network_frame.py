"""Drop-in ``TransformerNet`` of the per-frame mode (``use_Global=False``; reference:
``test/style_network_frame.py:361-394``).

Same 107-key ``state_dict`` as the global-mode class and the same kernels; what differs is the function
(SURVEY 3.2, quirk Q5): every InstanceNorm takes its statistics from the frame itself and never clamps
(:39-43), the six dynamic filters are predicted for every frame from content and style (:53-62), and
AdaIN(relu4_1) follows the filters directly (:339).  There is no ``add / compute / clean``: the reference
class has none either, and ``generate_real_video.py`` only calls them in global mode.
"""
from __future__ import annotations

from .style_network_global import TransformerNet as _GlobalNet


class TransformerNet(_GlobalNet):
    def add(self, patch):
        raise AttributeError("'TransformerNet' (frame mode) object has no attribute 'add'")

    def compute(self):
        raise AttributeError("'TransformerNet' (frame mode) object has no attribute 'compute'")

    def clean(self):
        raise AttributeError("'TransformerNet' (frame mode) object has no attribute 'clean'")

    def forward(self, input_frame):
        """[N,3,H,W] normalised RGB -> [N,3,H,W]: Decoder(Encoder(RGB2Gray(frame)), F_style), :392-394."""
        return self._eng().forward_frame(input_frame, kind=0, gray=True)

    def forward_u8(self, frame_bgr_u8):
        return self._eng().forward_frame(frame_bgr_u8, kind=1, gray=True)
