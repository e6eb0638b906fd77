# coding: utf-8
"""Stand-alone output samplers (the reference's mixture.py:118-156 and :221-270 entry points),
evaluated by libwn's CUDA sampler kernels over a whole (B,C,T) tensor.  The synthesis kernel has
the same arithmetic fused into its per-sample loop; these wrappers are for callers that sample
from a head-output tensor directly (e.g. on teacher-forced ``forward()`` output)."""
import ctypes as C

import torch

from . import _native as N


def _prep(y):
    if y.device.type != "cuda":
        raise RuntimeError("wavenet_vocoder_b200 samplers run on CUDA tensors only (no CPU fallback)")
    return y.detach().float().contiguous()


def _uniform(shape, device):
    return torch.empty(shape, device=device).uniform_(1e-5, 1.0 - 1e-5)


def sample_from_discretized_mix_logistic(y, log_scale_min=-7.0, clamp_log_scale=False, noise=None):
    """y: (B, 3K, T) -> (B, T) in [-1, 1].  ``noise``: optional dict u1 (T,B,K), u2 (T,B)."""
    y = _prep(y)
    B, O, T = y.shape
    assert O % 3 == 0
    K = O // 3
    if clamp_log_scale:
        y = y.clone()
        y[:, 2 * K:, :].clamp_(min=log_scale_min)
    u1 = noise["u1"].to(y.device).float().contiguous() if noise else _uniform((T, B, K), y.device)
    u2 = noise["u2"].to(y.device).float().contiguous() if noise else _uniform((T, B), y.device)
    out = torch.empty(B, T, device=y.device)
    st = torch.cuda.current_stream(y.device).cuda_stream
    N.check(N.lib().wn_sample_mol(y.data_ptr(), B, O, T, u1.data_ptr(), u2.data_ptr(), out.data_ptr(), st))
    return out


def sample_from_mix_gaussian(y, log_scale_min=-7.0, noise=None):
    """y: (B, 2 | 3K, T) -> (B, T) in [-1, 1].  ``noise``: optional dict z (T,B) [, u1 (T,B,K)]."""
    y = _prep(y)
    B, O, T = y.shape
    assert O == 2 or O % 3 == 0
    K = 1 if O == 2 else O // 3
    u1 = None
    if K > 1:
        u1 = noise["u1"].to(y.device).float().contiguous() if noise else _uniform((T, B, K), y.device)
    z = noise["z"].to(y.device).float().contiguous() if noise else torch.randn(T, B, device=y.device)
    out = torch.empty(B, T, device=y.device)
    st = torch.cuda.current_stream(y.device).cuda_stream
    N.check(N.lib().wn_sample_gauss(y.data_ptr(), B, O, T, u1.data_ptr() if u1 is not None else None,
                                    z.data_ptr(), out.data_ptr(), st))
    return out
