"""Drop-in ``warp`` and ``TemporalLoss`` (reference: ``train/loss_networks.py:20-38, 45-111``)."""
from __future__ import annotations

import random

import torch

from . import _lib as L


def _check(x, flo):
    if not (x.is_cuda and flo.is_cuda):
        raise RuntimeError("rerevst_b200.warp runs on CUDA tensors only")
    if x.dtype != torch.float32 or flo.dtype != torch.float32:
        raise TypeError("warp expects float32 tensors")
    B, C, H, W = x.shape
    if tuple(flo.shape) != (B, 2, H, W):
        raise ValueError(f"flow must be [{B},2,{H},{W}], got {tuple(flo.shape)}")
    return B, C, H, W


class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flo):
        B, C, H, W = _check(x, flo)
        x, flo = x.contiguous(), flo.contiguous()
        out = torch.empty_like(x)
        L.check(L.lib().rrv_warp_nearest_border(x.data_ptr(), flo.data_ptr(), B, C, H, W, out.data_ptr(), 0, L.stream()),
                "rrv_warp_nearest_border")
        ctx.save_for_backward(flo)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (flo,) = ctx.saved_tensors
        B, C, H, W = grad_out.shape
        grad_out = grad_out.contiguous()
        gx = torch.zeros(grad_out.shape, dtype=grad_out.dtype, device=grad_out.device)     # contiguous NCHW like the kernel writes
        L.check(L.lib().rrv_warp_backward(grad_out.data_ptr(), flo.data_ptr(), B, C, H, W, gx.data_ptr(),
                                          L.stream()), "rrv_warp_backward")
        return gx, None


def warp(x, flo, padding_mode="border"):
    """Nearest-neighbour, border-padded flow warp (loss_networks.py:20-38)."""
    if padding_mode != "border":
        raise NotImplementedError("only padding_mode='border' (the reference's default and only use)")
    return _Warp.apply(x, flo)


def warp_indices(flo):
    """int32 [B,H,W,2] = (iy, ix): the source pixel of every output pixel (bit-exact contract)."""
    B, _, H, W = flo.shape
    dummy = torch.zeros((B, 1, H, W), dtype=torch.float32, device=flo.device)
    out = torch.empty_like(dummy)
    idx = torch.empty((B, H, W, 2), dtype=torch.int32, device=flo.device)
    flo = flo.contiguous()
    L.check(L.lib().rrv_warp_nearest_border(dummy.data_ptr(), flo.data_ptr(), B, 1, H, W, out.data_ptr(),
                                            idx.data_ptr(), L.stream()), "rrv_warp_nearest_border")
    return idx


class TemporalLoss(torch.nn.Module):
    """The Compound Regularization loss (loss_networks.py:45-111).  ``forward`` is one fused kernel
    (warp + L1 mean) when no gradient is needed, and warp (with backward) + mean otherwise.
    Fake-flow synthesis (GenerateFakeFlow :71-86) stays on the host in numpy/cv2 like in the reference and draws
    from ``np.random`` / ``random`` in the reference's order, so equal seeds give the reference's flow."""

    def __init__(self, data_sigma=True, data_w=True, noise_level=0.001, motion_level=8, shift_level=10):
        super().__init__()
        self.data_sigma, self.data_w = data_sigma, data_w
        self.noise_level, self.motion_level, self.shift_level = noise_level, motion_level, shift_level

    def GaussianNoise(self, ins, mean=0, stddev=0.001):
        stddev = stddev + random.random() * stddev
        return ins + torch.empty_like(ins).normal_(mean, stddev)

    def GenerateFakeFlow(self, height, width):
        """(:71-86) [2, H, W] float32 on the host: low-resolution normal noise (sigma = motion_level px) resized to the frame,
        a random global shift of up to shift_level px per axis, 100 x 100 box blur."""
        import cv2
        import numpy as np
        if self.motion_level > 0:
            flow = np.random.normal(0, scale=self.motion_level, size=[height // 100, width // 100, 2])
            flow = cv2.resize(flow, (width, height))
            flow[:, :, 0] += random.randint(-self.shift_level, self.shift_level)
            flow[:, :, 1] += random.randint(-self.shift_level, self.shift_level)
            flow = cv2.blur(flow, (100, 100))
        else:
            flow = np.ones([width, height, 2])       # (sic: the reference builds this case as [W, H, 2], :82)
            flow[:, :, 0] = random.randint(-self.shift_level, self.shift_level)
            flow[:, :, 1] = random.randint(-self.shift_level, self.shift_level)
        return torch.from_numpy(flow.transpose((2, 0, 1))).float()

    def GenerateFakeData(self, first_frame, forward_flow=None):
        """(:88-104); the flow is synthesised on the host unless the caller supplies one ([2,H,W] or [B,2,H,W])."""
        if self.data_w:
            if forward_flow is None:
                forward_flow = self.GenerateFakeFlow(first_frame.shape[2], first_frame.shape[3])
            if forward_flow.dim() == 3:
                forward_flow = forward_flow.unsqueeze(0)
            forward_flow = forward_flow.to(first_frame.device).expand(first_frame.shape[0], 2, *first_frame.shape[2:]).contiguous()
            second = warp(first_frame, forward_flow)
        else:
            second, forward_flow = first_frame.clone(), None
        if self.data_sigma:
            second = self.GaussianNoise(second, stddev=self.noise_level)
        return second, forward_flow

    def forward(self, first_frame, second_frame, forward_flow):
        if not self.data_w:
            return torch.mean(torch.abs(first_frame - second_frame)), first_frame
        if torch.is_grad_enabled() and (first_frame.requires_grad or second_frame.requires_grad):
            w = warp(first_frame, forward_flow)
            return torch.mean(torch.abs(w - second_frame)), w
        B, C, H, W = _check(first_frame, forward_flow)
        if not second_frame.is_cuda or second_frame.dtype != torch.float32 or second_frame.shape != first_frame.shape:
            # the fused kernel reads second_frame element for element: anything else takes the reference's own route
            # (broadcasting / dtype promotion / the usual PyTorch errors)
            w = warp(first_frame, forward_flow)
            return torch.mean(torch.abs(w - second_frame)), w
        first, second, flo = first_frame.contiguous(), second_frame.contiguous(), forward_flow.contiguous()
        warped = torch.empty_like(first)
        acc = torch.empty(1, dtype=torch.float64, device=first.device)
        loss = torch.empty(1, dtype=torch.float32, device=first.device)
        L.check(L.lib().rrv_temporal_loss(first.data_ptr(), second.data_ptr(), flo.data_ptr(), B, C, H, W,
                                          warped.data_ptr(), acc.data_ptr(), loss.data_ptr(), L.stream()), "rrv_temporal_loss")
        return loss[0], warped
