# coding: utf-8
"""``WaveNet``: the reference's class surface (wavenet.py:63-361 there) over the B200 engine.

Same constructor keywords, same ``state_dict`` keys, same helper methods, same
``incremental_forward(initial_input, c=, g=, T=, test_inputs=, tqdm=, softmax=, quantize=,
log_scale_min=)`` signature and return layout, so ``synthesis.batch_wavegen`` / ``wavegen`` /
``train.eval_model`` of the reference work unchanged with this class.  What differs is what runs:
the per-sample loop is one persistent CUDA kernel launch (csrc/wn_kernel.cuh) instead of ~600
ATen calls per sample.

Conscious divergences from the reference (SURVEY.md 7 "quirks"):
  * ``incremental_forward`` needs the module on a CUDA device; on CPU it raises (no fallback).
  * speaker ids for a batch (``g`` of shape (B,1)) are embedded per row; the reference reshapes
    them with a stale B=1 (wavenet.py:265) and fails for B>1 unless test_inputs is given.
  * sampling noise comes from a Philox generator on the device, seeded from torch's global CPU
    generator (so ``torch.manual_seed`` still makes runs reproducible); pass ``noise=`` (the
    tensors the reference would have drawn, see oracle.predraw_noise) for bit-level comparisons.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import upsample
from .engine import SynthesisEngine
from .modules import Conv1d1x1, Embedding, ResidualConv1dGLU


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    """(kernel_size - 1) * sum(dilations) + 1, dilations cycling ``num_cycles`` times."""
    assert total_layers % num_cycles == 0
    per = total_layers // num_cycles
    return (kernel_size - 1) * sum(dilation(i % per) for i in range(total_layers)) + 1


def _expand_global_features(B, T, g, bct=True):
    if g is None:
        return None
    g = g.unsqueeze(-1) if g.dim() == 2 else g
    g = g.expand(B, -1, T)
    return g.contiguous() if bct else g.transpose(1, 2).contiguous()


class WaveNet(nn.Module):
    def __init__(self, out_channels=256, layers=20, stacks=2, residual_channels=512,
                 gate_channels=512, skip_out_channels=512, kernel_size=3, dropout=1 - 0.95,
                 cin_channels=-1, gin_channels=-1, n_speakers=None,
                 upsample_conditional_features=False, upsample_net="ConvInUpsampleNetwork",
                 upsample_params={"upsample_scales": [4, 4, 4, 4]}, scalar_input=False,
                 use_speaker_embedding=False, output_distribution="Logistic", cin_pad=0):
        super().__init__()
        assert layers % stacks == 0
        self.scalar_input = scalar_input
        self.out_channels = out_channels
        self.cin_channels = cin_channels
        self.gin_channels = gin_channels
        self.output_distribution = output_distribution
        self.layers, self.stacks, self.kernel_size = layers, stacks, kernel_size
        self.residual_channels, self.gate_channels = residual_channels, gate_channels
        self.skip_out_channels = skip_out_channels
        per_stack = layers // stacks
        self.first_conv = Conv1d1x1(1 if scalar_input else out_channels, residual_channels)
        self.conv_layers = nn.ModuleList([
            ResidualConv1dGLU(residual_channels, gate_channels, kernel_size=kernel_size,
                              skip_out_channels=skip_out_channels, bias=True,
                              dilation=2 ** (i % per_stack), dropout=dropout,
                              cin_channels=cin_channels, gin_channels=gin_channels)
            for i in range(layers)])
        self.last_conv_layers = nn.ModuleList([
            nn.ReLU(inplace=True), Conv1d1x1(skip_out_channels, skip_out_channels),
            nn.ReLU(inplace=True), Conv1d1x1(skip_out_channels, out_channels)])
        if gin_channels > 0 and use_speaker_embedding:
            assert n_speakers is not None
            self.embed_speakers = Embedding(n_speakers, gin_channels, padding_idx=None, std=0.1)
        else:
            self.embed_speakers = None
        if upsample_conditional_features:
            self.upsample_net = getattr(upsample, upsample_net)(**upsample_params)
        else:
            self.upsample_net = None
        self.receptive_field = receptive_field_size(layers, stacks, kernel_size)
        self._engine: Optional[SynthesisEngine] = None
        self._engine_key = None
        self._native_upsample = False

    # ------------------------------------------------------------------ small API of the reference
    def has_speaker_embedding(self):
        return self.embed_speakers is not None

    def local_conditioning_enabled(self):
        return self.cin_channels > 0

    def clear_buffer(self):
        """The engine's queues live only inside one synthesis call, so there is nothing to clear."""
        return None

    def make_generation_fast_(self):
        def strip(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                return
        self.apply(strip)

    # ------------------------------------------------------------------ teacher-forced batch forward
    def forward(self, x, c=None, g=None, softmax=False):
        """x: (B,C,T) -> (B,out_channels,T).  Plain PyTorch (training / likelihood path)."""
        B, _, T = x.size()
        if g is not None and self.embed_speakers is not None:
            g = self.embed_speakers(g.view(B, -1)).transpose(1, 2)
            assert g.dim() == 3
        g_bct = _expand_global_features(B, T, g, bct=True)
        if c is not None and self.upsample_net is not None:
            c = self.upsample_net(c)
            assert c.size(-1) == x.size(-1)
        h = self.first_conv(x)
        skips = 0
        for layer in self.conv_layers:
            h, s = layer(h, c, g_bct)
            skips = skips + s
        h = skips * math.sqrt(1.0 / len(self.conv_layers))
        for layer in self.last_conv_layers:
            h = layer(h)
        return F.softmax(h, dim=1) if softmax else h

    # ------------------------------------------------------------------ engine management
    def _param_version(self):
        """Key of the packed-weight cache: identity and in-place version of every parameter plus a
        data-dependent fingerprint (one fused norm over all parameters), so that updates through
        ``p.data.copy_()`` / ``p.data = ...`` (EMA swaps, weight surgery), which do not bump ``_version``,
        are seen too."""
        params = list(self.parameters())
        dev = params[0].device
        norms = torch.stack(torch._foreach_norm([p.detach() for p in params])).double()
        sums = torch.stack([p.detach().reshape(-1)[0] for p in params]).double()
        finger = torch.cat([norms, sums]).cpu().numpy().tobytes()
        return (str(dev), finger) + tuple((id(p), p._version) for p in params)

    def invalidate_engine(self):
        """Forget the packed weights (they are rebuilt by the next incremental_forward)."""
        self._engine_key = None

    def load_state_dict(self, *args, **kwargs):
        self._engine_key = None
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._engine_key = None
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        # the engine holds ctypes handles and device buffers: never pickled / deep-copied with the module
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_key"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_engine", "_engine_key") else copy.deepcopy(v, memo)
        return new

    def _get_engine(self) -> SynthesisEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("WaveNet.incremental_forward runs on a CUDA device only (B200 engine, no CPU "
                               "fallback); move the model with .to('cuda')")
        key = self._param_version()
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            self._engine = SynthesisEngine(
                layers=self.layers, stacks=self.stacks, residual_channels=self.residual_channels,
                gate_channels=self.gate_channels, skip_out_channels=self.skip_out_channels,
                out_channels=self.out_channels, kernel_size=self.kernel_size,
                cin_channels=self.cin_channels, gin_channels=self.gin_channels,
                scalar_input=self.scalar_input, output_distribution=self.output_distribution,
                device=dev)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.load_state_dict(self.state_dict())
            # the upsample network runs on the device too when libwn covers its configuration
            self._native_upsample = self._engine.load_upsampler(self.upsample_net)
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ the hot path
    @torch.no_grad()
    def incremental_forward(self, initial_input=None, c=None, g=None, T=100, test_inputs=None,
                            tqdm=lambda x: x, softmax=True, quantize=True, log_scale_min=-50.0,
                            noise: Optional[Dict[str, torch.Tensor]] = None, seed=None,
                            return_params=False):
        """Autoregressive synthesis; arguments and return value as the reference's
        ``WaveNet.incremental_forward`` (wavenet.py:215-343):

        initial_input (B,C,1)|(B,1,C); c (B,C',Tc) frames (upsampled here) or (B,C',T)/(B,T,C');
        g (B,)|(B,1) speaker ids or (B,gin[,1]) features; test_inputs (B,C,T')|(B,T',C) for teacher
        forcing.  Returns (B,1,T) for scalar input, else (B,out_channels,T).
        ``log_scale_min`` is accepted and, as in the reference (mixture.py:147-148 is never
        enabled by the caller), unused.  Extensions: ``noise`` (replayed draws), ``seed``,
        ``return_params`` (also return the per-step head outputs (B,O,T)).
        """
        if self.training:
            raise RuntimeError("incremental_forward only supports eval mode")       # conv.py:19-20
        eng = self._get_engine()
        dev = eng.device
        O = self.out_channels
        B = 1
        test_scalar = test_index = test_dense = None
        if test_inputs is not None:
            ti = test_inputs.to(dev)
            if self.scalar_input:
                if ti.size(1) == 1:
                    ti = ti.transpose(1, 2)
            elif ti.size(1) == O:
                ti = ti.transpose(1, 2)
            ti = ti.contiguous().float()                                            # (B,T',C)
            B = ti.size(0)
            T = ti.size(1) if T is None else max(int(T), ti.size(1))               # wavenet.py:255-258
            if self.scalar_input:
                test_scalar = ti.reshape(B, -1)
            else:
                idx = ti.argmax(-1)
                onehot = torch.zeros_like(ti).scatter_(-1, idx.unsqueeze(-1), 1.0)
                if torch.equal(onehot, ti):
                    test_index = idx.to(torch.int32)
                else:
                    test_dense = ti
        T = int(T)
        if c is not None:
            B = c.shape[0]
        g_vec = None
        if g is not None:
            g = g.to(dev)
            if self.embed_speakers is not None:
                g_vec = self.embed_speakers(g.view(g.size(0), -1).long())[:, 0, :]   # wavenet.py:263-266
            else:
                g_vec = g.reshape(g.size(0), -1).float()
            if g_vec.size(0) == 1 and B > 1:
                g_vec = g_vec.expand(B, -1)
            B = max(B, g_vec.size(0)) if c is None and test_inputs is None else B
        c_frames = None
        if c is not None:
            c = c.to(dev).float()
            if self.upsample_net is not None and getattr(self, "_native_upsample", False) \
                    and os.environ.get("WN_TORCH_UPSAMPLE", "0") != "1":
                assert c.dim() == 3 and c.size(1) == self.cin_channels
                assert eng.upsampled_length(c.size(-1)) == T                        # wavenet.py:276
                c_frames, c = c.contiguous(), None        # conv_in + stretch/smooth run inside libwn (csrc/wn_aux.cuh)
            else:
                if self.upsample_net is not None:
                    c = self.upsample_net(c)
                    assert c.size(-1) == T                                          # wavenet.py:276
                if c.size(-1) == T:
                    c = c.transpose(1, 2)
                c = c.contiguous()
                assert c.size(1) == T and c.size(2) == self.cin_channels
        initial = None
        initial_index = -1
        initial_rows = initial_dense = None
        if initial_input is not None and test_inputs is None:      # test_inputs override step 0 (wavenet.py:299-301)
            ii = initial_input.to(dev).float()
            if self.scalar_input:
                initial = ii.reshape(ii.size(0), -1)[:, 0].contiguous()
                if c is None and g is None:
                    B = initial.size(0)            # nothing else defines the batch (the reference assumes B=1 here)
                if initial.size(0) == 1 and B > 1:
                    initial = initial.expand(B).contiguous()
                if initial.size(0) != B:
                    raise ValueError("initial_input has %d rows but the batch is %d" % (initial.size(0), B))
            else:
                if ii.size(1) == O and ii.size(-1) != O:
                    ii = ii.transpose(1, 2)
                first = ii.reshape(ii.size(0), -1, O)[:, 0]                           # (B0, O), fed as is (wavenet.py:281-292)
                if c is None and g is None:
                    B = first.size(0)
                if first.size(0) == 1 and B > 1:
                    first = first.expand(B, -1)
                if first.size(0) != B:
                    raise ValueError("initial_input has %d rows but the batch is %d" % (first.size(0), B))
                idx = first.argmax(-1)
                onehot = torch.zeros_like(first).scatter_(-1, idx.unsqueeze(-1), 1.0)
                if torch.equal(onehot, first):
                    initial_rows = idx.to(torch.int32).contiguous()
                else:
                    initial_dense = first.contiguous()
        pinfo = eng.plan(B)
        if (self.scalar_input and B > pinfo["batch_tile"] and test_inputs is None and noise is None and not return_params
                and pinfo["engine"] == 5 and os.environ.get("WN_CONCURRENT_TILES", "1") != "0"):
            # more utterances than one batch tile, free running, device-drawn noise: two half-grid engines run two
            # tiles at the same time (engine.generate_concurrent; +44 % samples/s for 8 utterances on one B200)
            out = eng.generate_concurrent(B=B, T=T, c=c, c_frames=c_frames, g=g_vec, initial=initial, seed=seed, sync=False)
            bar = tqdm(range(T))
            eng.sync_concurrent()
            if hasattr(bar, "update"):
                bar.update(T)
            if hasattr(bar, "close"):
                bar.close()
            return out.view(B, 1, T)
        out, params = eng.generate(
            B=B, T=T, c=c, c_frames=c_frames, g=g_vec, initial=initial, initial_index=initial_index,
            initial_rows=initial_rows, initial_dense=initial_dense,
            test_scalar=test_scalar, test_index=test_index, test_dense=test_dense,
            softmax=bool(softmax), quantize=bool(quantize), noise=noise, seed=seed,
            want_params=return_params, sync=False)
        # progress-bar compatibility in O(1): the T steps are one kernel launch already in flight
        bar = tqdm(range(T))
        eng.sync()
        if hasattr(bar, "update"):
            bar.update(T)
        if hasattr(bar, "close"):
            bar.close()
        if self.scalar_input:
            y = out.view(B, 1, T)
        elif quantize:
            y = torch.zeros(B, O, T, device=dev).scatter_(1, out.long().unsqueeze(1), 1.0)
        else:
            y = out
        return (y, params) if return_params else y
