# coding: utf-8
"""Parameter containers of the dilated stack.

These modules exist for two reasons only: (1) they give the model the SAME ``state_dict`` keys
and shapes as the reference (``conv_layers.N.conv.weight_g`` ..., modules.py:13-39 and :52-110 in
the reference), so its checkpoints load unchanged; (2) they provide the teacher-forced batch
``forward`` used for training/likelihood and as a cross-check.  The per-sample incremental step
of the reference (modules.py:112-163, conv.py:17-46) is NOT implemented here: it lives in
csrc/wn_kernel.cuh and is reached through ``WaveNet.incremental_forward``.
"""
import math
import warnings

import torch
from torch import nn
from torch.nn import functional as F


def _normed_conv1d(cin, cout, ksize, dilation=1, padding=0, bias=True):
    conv = nn.Conv1d(cin, cout, ksize, padding=padding, dilation=dilation, bias=bias)
    nn.init.kaiming_normal_(conv.weight, nonlinearity="relu")
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)      # the old API keeps the reference's weight_g/weight_v keys
        return nn.utils.weight_norm(conv)


def Conv1d(in_channels, out_channels, kernel_size, dropout=0, **kwargs):
    return _normed_conv1d(in_channels, out_channels, kernel_size, **kwargs)


def Conv1d1x1(in_channels, out_channels, bias=True):
    return _normed_conv1d(in_channels, out_channels, 1, bias=bias)


def Embedding(num_embeddings, embedding_dim, padding_idx, std=0.01):
    emb = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    emb.weight.data.normal_(0, std)
    return emb


class ResidualConv1dGLU(nn.Module):
    """One gated residual layer (dilated causal conv -> +c +g -> tanh*sigmoid -> 1x1 skip/out)."""

    def __init__(self, residual_channels, gate_channels, kernel_size, skip_out_channels=None,
                 cin_channels=-1, gin_channels=-1, dropout=1 - 0.95, padding=None, dilation=1,
                 causal=True, bias=True):
        super().__init__()
        self.dropout = dropout
        self.causal = causal
        skip_out_channels = residual_channels if skip_out_channels is None else skip_out_channels
        if padding is None:
            padding = (kernel_size - 1) * dilation if causal else (kernel_size - 1) // 2 * dilation
        self.conv = _normed_conv1d(residual_channels, gate_channels, kernel_size, dilation=dilation,
                                   padding=padding, bias=bias)
        self.conv1x1c = Conv1d1x1(cin_channels, gate_channels, bias=False) if cin_channels > 0 else None
        self.conv1x1g = Conv1d1x1(gin_channels, gate_channels, bias=False) if gin_channels > 0 else None
        half = gate_channels // 2
        self.conv1x1_out = Conv1d1x1(half, residual_channels, bias=bias)
        self.conv1x1_skip = Conv1d1x1(half, skip_out_channels, bias=bias)

    def forward(self, x, c=None, g=None):
        """Batch (teacher-forced) form over a whole sequence. x: (B,R,T)."""
        T = x.size(-1)
        z = self.conv(F.dropout(x, p=self.dropout, training=self.training))
        if self.causal:
            z = z[:, :, :T]
        if c is not None:
            assert self.conv1x1c is not None
            z = z + self.conv1x1c(c)
        if g is not None:
            assert self.conv1x1g is not None
            z = z + self.conv1x1g(g)
        a, b = z.chunk(2, dim=1)
        y = torch.tanh(a) * torch.sigmoid(b)
        return (self.conv1x1_out(y) + x) * math.sqrt(0.5), self.conv1x1_skip(y)
