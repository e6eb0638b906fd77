# coding: utf-8
"""CPU oracle for the autoregressive WaveNet synthesis path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a flat weight dictionary, the algorithm that the
reference (r9y9/wavenet_vocoder) executes in ``WaveNet.incremental_forward``.  It exists so that
the CUDA path can be checked against something that runs where the reference itself is not
present (the GPU box has no ``/root/reference``).  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it, and only as the checker / the timed CPU baseline.

Pinning: ``tests/golden/make_golden.py`` (run in the build container, where ``/root/reference``
is importable) drives the unmodified reference and this oracle with identical weights, inputs and
torch RNG seed and asserts bit-equality of the per-step head outputs and of the sampled waveform
for every output head; the vectors it writes are committed under ``tests/golden/`` and re-checked
by ``tests/test_oracle_golden.py`` on every run.  Parity status: PINNED (bit-exact on this torch
build, torch 2.11 CPU).

All arithmetic is fp32 on torch's CPU kernels, in the same op order as the reference so that the
results are bit-identical to it (same ``F.linear`` shapes -> same BLAS path).

Reference citations are ``file:line`` under ``/root/reference``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration and weights
# --------------------------------------------------------------------------------------------
@dataclass
class PathConfig:
    """Shape of the dilated stack; field names follow ``wavenet.py:98-111``."""
    out_channels: int = 256
    layers: int = 20
    stacks: int = 2
    residual_channels: int = 512
    gate_channels: int = 512
    skip_out_channels: int = 512
    kernel_size: int = 3
    cin_channels: int = -1
    gin_channels: int = -1
    scalar_input: bool = False
    output_distribution: str = "Logistic"

    def dilations(self) -> List[int]:
        # wavenet.py:117-126  dilation = 2**(layer % layers_per_stack)
        assert self.layers % self.stacks == 0
        per = self.layers // self.stacks
        return [2 ** (i % per) for i in range(self.layers)]


def receptive_field_size(total_layers, num_cycles, kernel_size, dilation=lambda x: 2 ** x):
    """wavenet.py:42-60 — (kw-1)*sum(dilations)+1."""
    assert total_layers % num_cycles == 0
    per = total_layers // num_cycles
    return (kernel_size - 1) * sum(dilation(i % per) for i in range(total_layers)) + 1


def _fold(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """Weight of one conv with weight-norm folded (modules.py:13-18: w = g * v / ||v||, norm over
    all dims but 0).  Accepts the stripped form (``make_generation_fast_``, wavenet.py:355-361)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].detach().float()
    g = sd[prefix + ".weight_g"].detach().float()
    v = sd[prefix + ".weight_v"].detach().float()
    # torch._weight_norm(v, g, dim=0): v * (g / norm_except_dim(v, 2, 0))
    return torch._weight_norm(v, g, 0)


def weights_from_state_dict(cfg: PathConfig, sd: Dict[str, torch.Tensor]) -> Dict[str, object]:
    """Flatten a reference ``state_dict`` into the linearised matrices the incremental path uses.

    Conv weights (G, R, kw) are linearised tap-major as conv.py:51-62 does:
    ``W.transpose(1, 2).contiguous().view(G, kw*R)`` so column ``k*R + r`` is tap k (k=0 oldest).
    """
    def lin(prefix):
        w = _fold(sd, prefix)                      # (out, in, kw)
        out = w.size(0)
        return w.transpose(1, 2).contiguous().view(out, -1)

    def bias(prefix):
        b = sd.get(prefix + ".bias")
        return None if b is None else b.detach().float()

    w: Dict[str, object] = {
        "first_w": lin("first_conv"), "first_b": bias("first_conv"),
        "last_a_w": lin("last_conv_layers.1"), "last_a_b": bias("last_conv_layers.1"),
        "last_b_w": lin("last_conv_layers.3"), "last_b_b": bias("last_conv_layers.3"),
        "layers": [],
    }
    for i in range(cfg.layers):
        p = "conv_layers.%d." % i
        lay = {
            "conv_w": lin(p + "conv"), "conv_b": bias(p + "conv"),
            "out_w": lin(p + "conv1x1_out"), "out_b": bias(p + "conv1x1_out"),
            "skip_w": lin(p + "conv1x1_skip"), "skip_b": bias(p + "conv1x1_skip"),
            "c_w": lin(p + "conv1x1c") if cfg.cin_channels > 0 else None,
            "g_w": lin(p + "conv1x1g") if cfg.gin_channels > 0 else None,
        }
        w["layers"].append(lay)
    if "embed_speakers.weight" in sd:
        w["embed"] = sd["embed_speakers.weight"].detach().float()
    return w


# --------------------------------------------------------------------------------------------
# noise sources.  The reference draws from torch's global CPU generator in a fixed order; the
# "replay" source hands back pre-drawn tensors in that same order so a device kernel can consume
# identical noise (SURVEY.md 8(c) recipe 2).
# --------------------------------------------------------------------------------------------
class GlobalNoise:
    """Draw exactly as the reference does (mixture.py:138,151; Normal.sample; multinomial)."""

    def uniform(self, shape):
        return torch.empty(shape).uniform_(1e-5, 1.0 - 1e-5)

    def normal(self, shape):
        # torch.distributions.Normal.sample -> torch.normal(loc.expand(shape), scale.expand(shape))
        # which equals loc + scale * N(0,1) drawn with empty(shape).normal_()
        return torch.empty(shape).normal_()

    def exponential(self, shape):
        return torch.empty(shape).exponential_(1.0)


class ReplayNoise:
    """Hand back recorded per-step draws (lists of tensors, consumed in order)."""

    def __init__(self, uniform=None, normal=None, exponential=None):
        self._u = list(uniform or [])
        self._n = list(normal or [])
        self._e = list(exponential or [])

    def uniform(self, shape):
        t = self._u.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t

    def normal(self, shape):
        t = self._n.pop(0)
        assert tuple(t.shape) == tuple(shape)
        return t

    def exponential(self, shape):
        t = self._e.pop(0)
        assert tuple(t.shape) == tuple(shape)
        return t


def predraw_noise(cfg: PathConfig, B: int, T: int, seed: int):
    """Draw T steps of sampler noise in the reference's order/shape under ``manual_seed(seed)``.

    Returns a dict of stacked tensors (the layout the C-ABI replay mode takes):
      MoL      : u1 (T,B,K), u2 (T,B)          mixture.py:138, :151
      Normal   : z (T,B) [+ u1 (T,B,K) if O==3K]   mixture.py:247, :265-267
      softmax  : e (T,B,O)                     wavenet.py:334-335 (multinomial -> exponential_)
    """
    torch.manual_seed(seed)
    src = GlobalNoise()
    O = cfg.out_channels
    out: Dict[str, torch.Tensor] = {}
    if cfg.scalar_input:
        if cfg.output_distribution == "Logistic":
            K = O // 3
            u1, u2 = [], []
            for _ in range(T):
                u1.append(src.uniform((B, 1, K)))
                u2.append(src.uniform((B, 1)))
            out["u1"] = torch.stack(u1).view(T, B, K)
            out["u2"] = torch.stack(u2).view(T, B)
        else:
            K = 1 if O == 2 else O // 3
            u1, z = [], []
            for _ in range(T):
                if K > 1:
                    u1.append(src.uniform((B, 1, K)))
                z.append(src.normal((B, 1)))
            if K > 1:
                out["u1"] = torch.stack(u1).view(T, B, K)
            out["z"] = torch.stack(z).view(T, B)
    else:
        e = [src.exponential((B, O)) for _ in range(T)]
        out["e"] = torch.stack(e).view(T, B, O)
    return out


def replay_from_predrawn(cfg: PathConfig, noise: Dict[str, torch.Tensor]) -> ReplayNoise:
    u, n, e = [], [], []
    if "e" in noise:
        e = [noise["e"][t] for t in range(noise["e"].size(0))]
    else:
        T = (noise["u2"] if "u2" in noise else noise["z"]).size(0)
        for t in range(T):
            if "u1" in noise:
                u.append(noise["u1"][t].unsqueeze(1))          # (B,1,K)
            if "u2" in noise:
                u.append(noise["u2"][t].unsqueeze(1))          # (B,1)
            if "z" in noise:
                n.append(noise["z"][t].unsqueeze(1))           # (B,1)
    return ReplayNoise(uniform=u, normal=n, exponential=e)


# --------------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------------
def _gumbel_argmax(logits: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    # mixture.py:138-140
    return (logits - torch.log(-torch.log(u))).max(dim=-1)[1]


def _select(y: torch.Tensor, lo: int, hi: int, idx: torch.Tensor) -> torch.Tensor:
    # mixture.py:109-115,143-146: one-hot multiply then sum over the mixture axis
    one_hot = torch.zeros(idx.size() + (hi - lo,)).scatter_(idx.dim(), idx.unsqueeze(-1), 1.0)
    return torch.sum(y[..., lo:hi] * one_hot, dim=-1)


def sample_mol(y_b1o: torch.Tensor, noise) -> torch.Tensor:
    """mixture.py:118-156 on one step.  ``y_b1o`` is (B,1,3K) (already B x T x C).  Returns (B,1).
    ``log_scale_min`` is accepted by the reference but unused unless clamp_log_scale (never set
    by wavenet.py:324-325), so it does not appear here."""
    K = y_b1o.size(-1) // 3
    u1 = noise.uniform((y_b1o.size(0), 1, K))
    idx = _gumbel_argmax(y_b1o[:, :, :K], u1)
    means = _select(y_b1o, K, 2 * K, idx)
    log_scales = _select(y_b1o, 2 * K, 3 * K, idx)
    u2 = noise.uniform(tuple(means.size()))
    x = means + torch.exp(log_scales) * (torch.log(u2) - torch.log(1.0 - u2))
    return torch.clamp(torch.clamp(x, min=-1.0), max=1.0)


def sample_gaussian(y_b1o: torch.Tensor, noise) -> torch.Tensor:
    """mixture.py:221-270 on one step.  (B,1,2) single Gaussian or (B,1,3K) mixture -> (B,1)."""
    O = y_b1o.size(-1)
    if O == 2:
        means, log_scales = y_b1o[:, :, 0], y_b1o[:, :, 1]
    else:
        K = O // 3
        if K > 1:
            u1 = noise.uniform((y_b1o.size(0), 1, K))
            idx = _gumbel_argmax(y_b1o[:, :, :K], u1)
            means = _select(y_b1o, K, 2 * K, idx)
            log_scales = _select(y_b1o, 2 * K, 3 * K, idx)
        else:
            means, log_scales = y_b1o[:, :, 1], y_b1o[:, :, 2]
    scales = torch.exp(log_scales)
    z = noise.normal(tuple(means.size()))
    x = means + scales * z                         # == Normal(means, scales).sample() bit-exactly
    return torch.clamp(x, min=-1.0, max=1.0)


def sample_categorical(p_bo: torch.Tensor, noise) -> torch.Tensor:
    """wavenet.py:333-335: OneHotCategorical(p).sample().  Categorical renormalises ``probs`` and
    torch.multinomial(n=1) takes argmax(p / q), q ~ Exp(1).  Returns one-hot float (B,O)."""
    p = p_bo / p_bo.sum(-1, keepdim=True)
    q = noise.exponential(tuple(p.size()))
    idx = torch.argmax(p / q, dim=-1)
    return torch.zeros_like(p).scatter_(1, idx.unsqueeze(-1), 1.0)


# --------------------------------------------------------------------------------------------
# one layer / one step
# --------------------------------------------------------------------------------------------
class _Queue:
    """conv.py:32-44: shift register of the last (kw-1)*d+1 layer inputs, zero initialised
    (== causal zero padding), read with stride d."""

    def __init__(self, kw: int, d: int):
        self.kw, self.d, self.buf = kw, d, None

    def push_and_gather(self, x_b1r: torch.Tensor) -> torch.Tensor:
        B = x_b1r.size(0)
        if self.kw == 1:
            return x_b1r.reshape(B, -1)
        if self.buf is None:
            self.buf = x_b1r.new_zeros(B, self.kw + (self.kw - 1) * (self.d - 1), x_b1r.size(2))
        else:
            self.buf[:, :-1, :] = self.buf[:, 1:, :].clone()
        self.buf[:, -1, :] = x_b1r[:, -1, :]
        taps = self.buf if self.d == 1 else self.buf[:, 0::self.d, :].contiguous()
        return taps.reshape(B, -1)


def _layer_step(lay, q: _Queue, x, ct, gt):
    """modules.py:127-163 in incremental mode.  x (B,1,R) -> (x', skip)."""
    B = x.size(0)
    residual = x
    z = F.linear(q.push_and_gather(x), lay["conv_w"], lay["conv_b"]).view(B, 1, -1)   # conv.py:45
    a, b = z.split(z.size(-1) // 2, dim=-1)                                            # modules.py:138
    if ct is not None:
        cz = F.linear(ct.reshape(B, -1), lay["c_w"]).view(B, 1, -1)                    # modules.py:141-145
        ca, cb = cz.split(cz.size(-1) // 2, dim=-1)
        a, b = a + ca, b + cb
    if gt is not None:
        gz = F.linear(gt.reshape(B, -1), lay["g_w"]).view(B, 1, -1)                    # modules.py:148-152
        ga, gb = gz.split(gz.size(-1) // 2, dim=-1)
        a, b = a + ga, b + gb
    y = torch.tanh(a) * torch.sigmoid(b)                                               # modules.py:154
    s = F.linear(y.view(B, -1), lay["skip_w"], lay["skip_b"]).view(B, 1, -1)           # modules.py:157
    o = F.linear(y.view(B, -1), lay["out_w"], lay["out_b"]).view(B, 1, -1)             # modules.py:160
    return (o + residual) * math.sqrt(0.5), s                                          # modules.py:162


def incremental_forward(cfg: PathConfig, w: Dict[str, object],
                        initial_input: Optional[torch.Tensor] = None,
                        c: Optional[torch.Tensor] = None,
                        g: Optional[torch.Tensor] = None,
                        T: int = 100,
                        test_inputs: Optional[torch.Tensor] = None,
                        softmax: bool = True, quantize: bool = True,
                        noise=None,
                        params_out: Optional[list] = None,
                        progress: Callable = lambda x: x) -> torch.Tensor:
    """wavenet.py:215-343 restated.

    c: local conditioning ALREADY at sample rate, (B,C,T) or (B,T,C) (the upsample network runs
       before the loop, wavenet.py:272-278, and is outside this path).
    g: global conditioning as a float vector (B,gin) / (B,gin,1) (the embedding lookup,
       wavenet.py:263-268, is done by the caller; ``embed_speaker`` below does it).
    params_out: if a list, the per-step head output (B,O) is appended (the sampler input the
       reference never returns for scalar-input models).
    Returns (B,C,T) like the reference: C=1 for scalar input, else out_channels.
    """
    noise = noise or GlobalNoise()
    O = cfg.out_channels
    B = 1
    if test_inputs is not None:                                                   # wavenet.py:247-258
        if cfg.scalar_input:
            if test_inputs.size(1) == 1:
                test_inputs = test_inputs.transpose(1, 2).contiguous()
        elif test_inputs.size(1) == O:
            test_inputs = test_inputs.transpose(1, 2).contiguous()
        B = test_inputs.size(0)
        T = test_inputs.size(1) if T is None else max(T, test_inputs.size(1))
    T = int(T)
    if g is not None:
        g = g.reshape(g.size(0), -1)                                              # (B,gin)
    if c is not None:                                                             # wavenet.py:272-278
        B = c.shape[0]
        if c.size(-1) == T:
            c = c.transpose(1, 2).contiguous()
    if initial_input is None:                                                     # wavenet.py:281-292
        if cfg.scalar_input:
            initial_input = torch.zeros(B, 1, 1)
        else:
            initial_input = torch.zeros(B, 1, O)
            initial_input[:, :, 127] = 1
    elif initial_input.size(1) == O:
        initial_input = initial_input.transpose(1, 2).contiguous()

    dil = cfg.dilations()
    queues = [_Queue(cfg.kernel_size, d) for d in dil]
    L = cfg.layers
    outputs: List[torch.Tensor] = []
    current = initial_input
    for t in progress(range(T)):                                                  # wavenet.py:296
        if test_inputs is not None and t < test_inputs.size(1):
            current = test_inputs[:, t, :].unsqueeze(1)
        elif t > 0:
            current = outputs[-1]
        ct = None if c is None else c[:, t, :].unsqueeze(1)
        gt = None if g is None else g.unsqueeze(1)
        x = current.reshape(B, 1, -1)
        x = F.linear(x.view(B, -1), w["first_w"], w["first_b"]).view(B, 1, -1)    # wavenet.py:308
        skips = 0
        for lay, q in zip(w["layers"], queues):                                   # wavenet.py:310-312
            x, h = _layer_step(lay, q, x, ct, gt)
            skips = skips + h if isinstance(skips, int) else skips.add_(h)
        skips = skips * math.sqrt(1.0 / L)                                        # wavenet.py:313
        x = F.relu(skips)
        x = F.linear(x.view(B, -1), w["last_a_w"], w["last_a_b"])
        x = F.relu(x)
        x = F.linear(x, w["last_b_w"], w["last_b_b"]).view(B, 1, -1)              # wavenet.py:315-319
        if params_out is not None:
            params_out.append(x.view(B, -1).clone())
        if cfg.scalar_input:                                                      # wavenet.py:322-330
            if cfg.output_distribution == "Logistic":
                x = sample_mol(x, noise)
            elif cfg.output_distribution == "Normal":
                x = sample_gaussian(x, noise)
            else:
                raise AssertionError(cfg.output_distribution)
        else:                                                                     # wavenet.py:331-335
            x = F.softmax(x.view(B, -1), dim=1) if softmax else x.view(B, -1)
            if quantize:
                x = sample_categorical(x, noise)
        outputs.append(x)
    out = torch.stack(outputs)                                                    # T x B x C
    return out.transpose(0, 1).transpose(1, 2).contiguous()                       # B x C x T


def embed_speaker(w: Dict[str, object], g_ids: torch.Tensor) -> torch.Tensor:
    """wavenet.py:263-266: speaker id (B,)/(B,1) -> (B,gin)."""
    return F.embedding(g_ids.view(g_ids.size(0), -1), w["embed"])[:, 0, :]
