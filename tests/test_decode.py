# coding: utf-8
"""Decode step after the path: the device kernel (wn_decode, csrc/wn_aux.cuh) against the
numpy oracle (oracle/decode_oracle.py), and the oracle against the closed forms / scipy it restates.

Bar: bit-exact int16 for the "raw" input type (every operation is a single correctly rounded IEEE float32
operation on both sides, incl. the serial inv_preemphasis recursion); for the two mu-law types the only
non-IEEE step is powf, whose CUDA and libm implementations may differ in the last ulp, so int16 may differ by
one LSB there (asserted <= 1 LSB, and the fraction of differing samples reported by the assert message)."""
import numpy as np
import pytest
import torch

from oracle import decode_oracle as dorc


def test_oracle_matches_scipy_lfilter_float32():
    from scipy import signal
    rng = np.random.RandomState(0)
    x = (rng.randn(5000) * 0.3).astype(np.float32)
    ref = signal.lfilter(np.array([1.0], np.float32), np.array([1.0, -0.85], np.float32), x)
    assert ref.dtype == np.float32
    assert np.array_equal(dorc.inv_preemphasis(x, 0.85), ref)           # nnmnkwii: b, a cast to x.dtype
    e = dorc.inv_preemphasis(np.array([1.0, 0.0, 0.0], np.float32), 0.5)
    assert np.allclose(e, [1.0, 0.5, 0.25])


def test_oracle_mulaw_identities():
    x = np.linspace(-1, 1, 1001).astype(np.float32)
    mu = 255
    y = (np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)).astype(np.float32)     # nnmnkwii mulaw
    assert np.abs(dorc.inv_mulaw(y, mu) - x).max() < 1e-5
    q = ((y + 1) / 2 * mu).astype(np.int64)                                            # nnmnkwii mulaw_quantize
    assert np.abs(dorc.inv_mulaw_quantize(q, mu) - x).max() < 0.045          # one (truncated) bin at full scale
    assert dorc.inv_mulaw_quantize(np.array([0, 255]), mu).tolist() == [-1.0, 1.0]
    assert dorc.to_int16(np.array([-1.0, 0.0, 1.0, 0.99999], np.float32)).tolist() == [-32767, 0, 32767, 32766]


def test_decode_needs_a_gpu_tensor():
    from wavenet_vocoder_b200 import dispatch as D
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.decode_device(torch.zeros(1, 1, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("post,gain", [(None, 0.0), ("inv_preemphasis", 0.0), ("inv_preemphasis", 0.55), (None, 1.0)])
def test_raw_decode_is_bit_exact(post, gain):
    from wavenet_vocoder_b200 import dispatch as D
    rng = np.random.RandomState(1)
    B, T = 3, 5000
    y = np.clip(rng.randn(B, T) * 0.4, -1, 1).astype(np.float32)
    lens = [T, 1234, 4097]
    coef = 0.85 if post else 0.0
    f_ref, p_ref = dorc.decode(y, lens, "raw", 65536, coef, gain)
    pcm, flt = D.decode_device(torch.from_numpy(y).view(B, 1, T).cuda(), lens, "raw", 65536, post, gain, want_float=True)
    assert pcm.dtype == torch.int16 and tuple(pcm.shape) == (B, T)
    assert np.array_equal(flt.cpu().numpy(), f_ref)
    assert np.array_equal(pcm.cpu().numpy(), p_ref)


@pytest.mark.gpu
def test_long_recursion_carries_state_across_chunks():
    from wavenet_vocoder_b200 import dispatch as D
    rng = np.random.RandomState(2)
    y = (rng.randn(1, 240000) * 0.1).astype(np.float32)              # config 5 length: 235 chunks of the kernel
    _, p_ref = dorc.decode(y, None, "raw", 65536, 0.85, 0.0)
    pcm = D.decode_device(torch.from_numpy(y).view(1, 1, -1).cuda(), None, "raw", 65536, "inv_preemphasis", 0.0)
    assert np.array_equal(pcm.cpu().numpy(), p_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mulaw", "mulaw-quantize"])
def test_mulaw_decode_within_one_lsb(kind):
    from wavenet_vocoder_b200 import dispatch as D
    rng = np.random.RandomState(3)
    B, T, Q = 2, 4000, 256
    if kind == "mulaw-quantize":
        idx = rng.randint(0, Q, size=(B, T))
        y_hat = torch.zeros(B, Q, T).scatter_(1, torch.from_numpy(idx).unsqueeze(1), 1.0)
        f_ref, p_ref = dorc.decode(idx, None, kind, Q, 0.85, 0.0)
    else:
        y = rng.uniform(-1, 1, size=(B, T)).astype(np.float32)
        y_hat = torch.from_numpy(y).view(B, 1, T)
        f_ref, p_ref = dorc.decode(y, None, kind, Q, 0.85, 0.0)
    pcm, flt = D.decode_device(y_hat.cuda(), None, kind, Q, "inv_preemphasis", 0.0, want_float=True)
    d = np.abs(pcm.cpu().numpy().astype(np.int32) - p_ref.astype(np.int32))
    assert d.max() <= 1, (d.max(), float((d > 0).mean()))
    assert np.abs(flt.cpu().numpy() - f_ref).max() <= 2e-6
