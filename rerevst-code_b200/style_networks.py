"""Inference-side pieces of the training model (reference: ``train/style_networks.py``) that sit on the
temporal-loss path of ``train/train.py:375-388``: the ``Vgg19`` loss network,
``calc_mean_std``, ``TransformerNet.validation / style_loss / content_loss``.  The frozen ``Vgg19`` loss network is
differentiable w.r.t. its input (data-gradient convolutions on the same tensor-core kernel, SURVEY 8f N2 first half), so
``content_loss`` / ``style_loss`` / ``TemporalLoss`` gradients reach the styled frame; the stylizer's own backward (weight
gradients, per-frame InstanceNorm) is not built.

All arithmetic runs in ``librerevst_b200.so``; the MSE of two small [N,C] tables is the only torch op.
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn as nn

from .style_network_global import TransformerNet as _GlobalNet

vgg_outputs = namedtuple("VggOutputs", ["relu1_1", "relu2_1", "relu3_1", "relu4_1"])


def calc_mean_std(feat, eps=1e-5, _engine=None):
    """Per-sample channel mean and sqrt(unbiased var + eps) of an NCHW feature map (:95-103) -> ([N,C,1,1], [N,C,1,1]).
    ``feat`` must come from :meth:`Vgg19.forward` / :meth:`TransformerNet.vgg19` (NCHW views of NHWC device tensors)."""
    if eps != 1e-5:
        raise NotImplementedError("only the reference's eps=1e-5")
    if _engine is None:
        raise RuntimeError("calc_mean_std needs the engine that produced the features (use TransformerNet.calc_mean_std)")
    mean, std = _engine.feature_mean_std(feat)
    n, c = mean.shape
    return mean.view(n, c, 1, 1), std.view(n, c, 1, 1)


class _Vgg19Fn(torch.autograd.Function):
    """Vgg19.forward with a backward: the loss network is frozen (requires_grad False, train.py:166-175 optimises style_net only),
    so Loss.backward() needs its data gradients only -- engine.vgg_backward (tensor-core data-gradient convolutions)."""

    @staticmethod
    def forward(ctx, x, eng):
        feats, saved = eng.vgg_features_train(x, "Vgg19")
        ctx.eng, ctx.saved = eng, saved
        return feats

    @staticmethod
    def backward(ctx, g1, g2, g3, g4):
        return ctx.eng.vgg_backward(ctx.saved, (g1, g2, g3, g4), "Vgg19"), None


class TransformerNet(_GlobalNet):
    """The training-time model's evaluation API.  ``Vgg19.*`` weights are kept (the test-time class deletes them after
    the first style); ``validation`` is the frame-mode network without RGB2Gray (:556-559)."""

    def generate_style_features(self, style):
        self._eng().generate_style_features(style)
        self.F_style = self._engine.F_style

    def validation(self, cur_frame, style):
        eng = self._eng()
        eng.generate_style_features(style)
        self.F_style = eng.F_style
        return eng.forward_frame(cur_frame, kind=0, gray=False)

    def forward(self, input_frame):
        return self._eng().forward_frame(input_frame, kind=0, gray=True)

    def vgg19(self, x):
        """Vgg19.forward (:284-314): relu1_1 .. relu4_1 of a batch of normalised RGB images.  Differentiable w.r.t. ``x`` (the styled
        frame, train.py:376-414): under autograd the forward keeps the ReLU outputs and the backward runs on the same kernels."""
        if torch.is_grad_enabled() and x.requires_grad:
            return vgg_outputs(*_Vgg19Fn.apply(x, self._eng()))
        return vgg_outputs(*self._eng().vgg_features(x, "Vgg19"))

    def calc_mean_std(self, feat):
        if torch.is_grad_enabled() and feat.requires_grad:         # loss algebra on small reductions: autograd's own ops (:95-103)
            n, c = feat.shape[:2]
            flat = feat.reshape(n, c, -1)
            return flat.mean(dim=2).view(n, c, 1, 1), (flat.var(dim=2) + 1e-5).sqrt().view(n, c, 1, 1)
        return calc_mean_std(feat, _engine=self._eng())

    def style_loss(self, features_coded_Image, features_style):
        """:503-512: sum over the four levels of MSE(mean) + MSE(std)."""
        loss = 0.0
        for ft_x, ft_s in zip(features_coded_Image, features_style):
            mean_x, std_x = self.calc_mean_std(ft_x)
            mean_s, std_s = self.calc_mean_std(ft_s)
            loss = loss + nn.functional.mse_loss(mean_x, mean_s) + nn.functional.mse_loss(std_x, std_s)
        return loss

    def content_loss(self, features_coded_Image, features_content):
        """:514-516: MSE of the relu4_1 maps."""
        return nn.functional.mse_loss(features_coded_Image.relu4_1, features_content.relu4_1)
