# coding: utf-8
"""Local-conditioning upsamplers (frame rate -> sample rate).

They run ONCE per utterance before the sample loop (reference wavenet.py:272-276) and are outside
the hot path (SURVEY.md 8(f-1) lists them as the next row); they stay plain PyTorch here and keep
the reference's parameter names (upsample.py:29-85 there) so checkpoints load.
"""
import warnings

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F


class Stretch2d(nn.Module):
    def __init__(self, x_scale, y_scale, mode="nearest"):
        super().__init__()
        self.x_scale, self.y_scale, self.mode = x_scale, y_scale, mode

    def forward(self, x):
        return F.interpolate(x, scale_factor=(self.y_scale, self.x_scale), mode=self.mode)


class UpsampleNetwork(nn.Module):
    """Per scale s: nearest-neighbour stretch by s, then a (1 x 2s+1) smoothing conv."""

    def __init__(self, upsample_scales, upsample_activation="none", upsample_activation_params={},
                 mode="nearest", freq_axis_kernel_size=1, cin_pad=0, cin_channels=80):
        super().__init__()
        self.up_layers = nn.ModuleList()
        self.indent = cin_pad * int(np.prod(upsample_scales))
        for s in upsample_scales:
            ksize = (freq_axis_kernel_size, 2 * s + 1)
            conv = nn.Conv2d(1, 1, kernel_size=ksize, padding=((freq_axis_kernel_size - 1) // 2, s),
                             bias=False)
            conv.weight.data.fill_(1.0 / np.prod(ksize))
            self.up_layers.append(Stretch2d(s, 1, mode))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", FutureWarning)
                self.up_layers.append(nn.utils.weight_norm(conv))
            if upsample_activation != "none":
                self.up_layers.append(getattr(nn, upsample_activation)(**upsample_activation_params))

    def forward(self, c):
        c = c.unsqueeze(1)
        for layer in self.up_layers:
            c = layer(c)
        c = c.squeeze(1)
        return c[:, :, self.indent:-self.indent] if self.indent > 0 else c


class ConvInUpsampleNetwork(nn.Module):
    """A (2*cin_pad+1)-tap conv over frames for context, then UpsampleNetwork."""

    def __init__(self, upsample_scales, upsample_activation="none", upsample_activation_params={},
                 mode="nearest", freq_axis_kernel_size=1, cin_pad=0, cin_channels=80):
        super().__init__()
        self.conv_in = nn.Conv1d(cin_channels, cin_channels, kernel_size=2 * cin_pad + 1, bias=False)
        self.upsample = UpsampleNetwork(upsample_scales, upsample_activation, upsample_activation_params,
                                        mode, freq_axis_kernel_size, cin_pad=0, cin_channels=cin_channels)

    def forward(self, c):
        return self.upsample(self.conv_in(c))
