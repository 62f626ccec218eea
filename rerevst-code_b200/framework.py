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
        self.use_Global = bool(use_Global)
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
            while len(self._pin) >= 4:          # bounded: streams of varying resolution do not grow pinned memory forever
                self._pin.pop(next(iter(self._pin)))
            self._pin[key] = (torch.empty((1,) + key, dtype=torch.uint8).pin_memory(), torch.cuda.Event())
        pinned, copied = self._pin[key]
        copied.synchronize()                    # the previous asynchronous upload from this staging buffer has finished
        pinned[0].copy_(torch.from_numpy(img))
        dev = pinned.to(self.device, non_blocking=True)
        copied.record(torch.cuda.current_stream(self.device))
        return dev

    # ===== Sequence-Level Global Feature Sharing =====
    def add(self, patch):
        if not self.use_Global:
            return self.model.add(patch)        # AttributeError, as in the reference (frame mode has no pre-pass)
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

    def transfer(self, frame, crop=None, out_dtype="f32"):
        """frame: uint8 BGR HWC -> float32 BGR HWC in [0,255] (test/framework.py:106-118).
        crop=(y0, x0, h, w) additionally applies generate_real_video.py:167 on the device.
        out_dtype="u8" returns the frame rounded to uint8 the way cv2.imwrite stores a float32 image (:170): a quarter of
        the bytes to download."""
        out = self.transfer_device(frame, crop, out_dtype)
        return out.cpu().numpy()[0]

    @staticmethod
    def output_size(H, W):
        """Size of the stylized frame for an H x W input: (H // 8) * 8 x (W // 8) * 8 (three floor max-pools, three x2
        upsamples), e.g. 436 x 1024 -> 432 x 1024 exactly like the reference."""
        return (H // 8) * 8, (W // 8) * 8

    def transfer_stream(self, frames, crop=None, depth=4, pad_to=None, copy=True, out_dtype="f32", lanes=None, device_sink=None):
        """Generator over an iterable of uint8 BGR frames of one size: yields exactly what
        ``transfer(frame, crop)`` returns for each, in order, but pipelined -- the pinned-memory
        upload of frame i+1 (copy-in stream) and the download of frame i-1 (copy-out stream) overlap the
        kernels of frame i (current stream).  ``depth`` frames are in flight.

        ``copy=False`` yields a view of the pinned download buffer instead of a fresh array: valid until the generator is
        advanced again (enough for a consumer that writes the frame out before asking for the next one; saves a 25 MB host
        copy per 1080p frame, which is what limits 8 processes sharing one host).

        ``pad_to=(PH, PW)``: the frames are RAW; ReshapeTool.process (generate_real_video.py:66-83, reflect border of 64
        pixels up to PH x PW) runs on the device, and ``crop`` defaults to the raw frame's window (:167).

        ``out_dtype="u8"``: uint8 frames (see transfer): 6.2 MB instead of 24.9 MB per 1080p frame over PCIe.

        ``lanes``: frames whose kernels run concurrently (default 2 in global mode, where a frame is one graph replay without
        shared scratch; 1 in frame mode): consecutive frames alternate between two compute streams, so that the SMs one
        frame's layer leaves idle in its last persistent round run the other frame's kernels (engine.forward_graphed).

        ``device_sink(i, dev_frame, stream)``: called for every frame with the finished frame still on the device ([h, w, 3] tensor
        of ``out_dtype``, valid until the frame is yielded) and the copy-out stream current, e.g. to enqueue a GPU JPEG encode
        (video_io.MjpgWriter.encode); if it returns a callable, that is run right before the frame is yielded."""
        if out_dtype not in ("f32", "u8"):
            raise ValueError("out_dtype must be 'f32' or 'u8'")
        t_out = torch.uint8 if out_dtype == "u8" else torch.float32
        eng = self.model._eng()
        if lanes is None:
            lanes = 2 if self.use_Global else 1
        if lanes < 1 or (lanes > 1 and not self.use_Global):
            raise ValueError("lanes must be >= 1 (and 1 in frame mode)")
        depth = max(depth, lanes + 1)
        comp = eng.lane_streams(lanes)
        s_in, s_out = self._side_streams()
        slots, pending = [], []

        def finish(slot):
            slot["ev_out"].synchronize()
            fin, slot["fin"] = slot.get("fin"), None
            if fin is not None:
                fin()
            out = slot["host_out"].numpy()[0]
            return out.copy() if copy else out

        for i, frame in enumerate(frames):
            frame = np.ascontiguousarray(frame)
            if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
                raise ValueError("expected a uint8 HxWx3 BGR image (cv2.imread layout)")
            rH, rW = frame.shape[:2]                       # as uploaded
            H, W = pad_to if pad_to is not None else (rH, rW)      # as seen by the network
            oH, oW = self.output_size(H, W)                         # as returned by it
            if crop is not None:
                y0, x0, h, w = crop
            elif pad_to is not None:
                y0, x0, h, w = 64, 64, rH, rW
            else:
                y0, x0, h, w = 0, 0, oH, oW
            if y0 < 0 or x0 < 0 or y0 + h > oH or x0 + w > oW:
                raise ValueError(f"crop {(y0, x0, h, w)} outside the {oH}x{oW} stylized frame")
            if len(slots) < depth:
                slots.append(dict(host_in=torch.empty((1, rH, rW, 3), dtype=torch.uint8).pin_memory(),
                                  dev_in=torch.empty((1, rH, rW, 3), dtype=torch.uint8, device=self.device),
                                  dev_pad=(torch.empty((1, H, W, 3), dtype=torch.uint8, device=self.device)
                                           if pad_to is not None else None),
                                  dev_out=torch.empty((1, h, w, 3), dtype=t_out, device=self.device),
                                  host_out=torch.empty((1, h, w, 3), dtype=t_out).pin_memory(),
                                  ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(),
                                  ev_free=torch.cuda.Event()))
            slot = slots[i % depth]
            if len(pending) == depth:                       # this slot's previous frame must be handed out first
                yield finish(pending.pop(0))
            if tuple(slot["host_in"].shape[1:3]) != (rH, rW):
                raise ValueError("transfer_stream needs frames of one size")
            slot["host_in"][0].copy_(torch.from_numpy(frame))
            with torch.cuda.stream(s_in):
                s_in.wait_event(slot["ev_free"])            # the kernels that read dev_in last time are done
                slot["dev_in"].copy_(slot["host_in"], non_blocking=True)
                slot["ev_in"].record(s_in)
            lane = i % lanes
            cur = comp[lane]
            with torch.cuda.stream(cur):
                cur.wait_event(slot["ev_in"])
                net_in = slot["dev_in"]
                if pad_to is not None:
                    net_in = slot["dev_pad"]
                    L.check(L.lib().rrv_reflect_pad_u8(slot["dev_in"].data_ptr(), 1, rH, rW, 64, 64, H, W, net_in.data_ptr(), L.stream()),
                            "rrv_reflect_pad_u8")
                cur.wait_event(slot["ev_out"])                  # dev_out of this slot has been downloaded
                # the network's output is the lane's captured graph's static buffer: a device-to-device copy (~10 us at 1080p) into
                # this slot's buffer lets the lane's next frame start while this frame is still being downloaded
                slot["dev_out"].copy_(self._net(eng, net_in, (out_dtype, (y0, x0, h, w)), lane), non_blocking=True)
                slot["ev_free"].record(cur)
                slot["ev_done"].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(slot["ev_done"])
                slot["host_out"].copy_(slot["dev_out"], non_blocking=True)
                if device_sink is not None:
                    slot["fin"] = device_sink(i, slot["dev_out"][0], s_out)
                slot["ev_out"].record(s_out)
            pending.append(slot)
        while pending:
            yield finish(pending.pop(0))

    def _net(self, eng, dev_u8, post, lane=0):
        """uint8 NHWC frame on the device -> the finished [N, h, w, 3] BGR frame on the device.  Global mode: one CUDA-graph
        replay whose last kernel (the RGB head) also de-normalises, clamps, crops and converts; the returned tensor is the
        graph's static output, valid until the next call."""
        if self.use_Global:
            return eng.forward_graphed(dev_u8, kind=1, post=post, lane=lane)
        y = eng.forward_frame_graphed(dev_u8, kind=1, gray=True)
        return eng.postprocess(y, post[1], post[0])

    def _side_streams(self):
        if not hasattr(self, "_streams"):
            self._streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return self._streams

    def transfer_device(self, frame, crop=None, out_dtype="f32"):
        if out_dtype not in ("f32", "u8"):
            raise ValueError("out_dtype must be 'f32' or 'u8'")
        eng = self.model._eng()
        dev = self._upload(frame)
        oH, oW = self.output_size(dev.shape[1], dev.shape[2])
        crop = tuple(crop) if crop is not None else (0, 0, oH, oW)
        if crop[0] < 0 or crop[1] < 0 or crop[0] + crop[2] > oH or crop[1] + crop[3] > oW:
            raise ValueError(f"crop {crop} outside the {oH}x{oW} stylized frame")
        return self._net(eng, dev, (out_dtype, crop)).clone()       # the graph's static output is reused by the next call
