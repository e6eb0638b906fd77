# coding: utf-8
"""CPU oracle for the decode step that follows the synthesis path.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy float32, what the reference does to the model output before writing a wav file:

* ``synthesis.py:66-84`` (``batch_wavegen``): class ids -> ``P.inv_mulaw_quantize`` / companded scalars ->
  ``P.inv_mulaw`` / raw scalars untouched; then ``audio.inv_preemphasis`` when ``hparams.postprocess`` names it
  (``audio.py:57-58``); then division by ``hparams.global_gain_scale``;
* ``evaluate.py:215,247`` trim to the utterance's own length and clip to [-1, 1];
* ``evaluate.py:43-48`` ``to_int16``: ``(x * 32767).astype(np.int16)``.

``P`` is ``nnmnkwii.preprocessing`` (setup.py:23 pins ``nnmnkwii >= 0.0.11``), which is not installed here and is not
part of /root/reference, so its published algorithm is restated (nnmnkwii/preprocessing/generic.py):

    inv_mulaw(y, mu)          = sign(y) * (1.0 / mu) * ((1.0 + mu) ** abs(y) - 1.0)
    inv_mulaw_quantize(y, mu) = inv_mulaw(2 * y.astype(float32) / mu - 1, mu)
    inv_preemphasis(x, coef)  = scipy.signal.lfilter([1], [1, -coef], x)   with b, a cast to x.dtype

Parity status of this file: the arithmetic is pinned against scipy.signal.lfilter (float32) and against the
closed forms above by tests/test_decode.py; the nnmnkwii functions themselves cannot be run here ("parity
unpinned" for the two mu-law variants beyond those identities).
"""
import numpy as np


def inv_mulaw(y, mu=255):
    y = np.asarray(y, dtype=np.float32)
    return (np.sign(y) * np.float32(1.0 / mu) * (np.float32(1.0 + mu) ** np.abs(y) - np.float32(1.0))).astype(np.float32)


def inv_mulaw_quantize(idx, mu=255):
    y = np.float32(2.0) * np.asarray(idx).astype(np.float32) / np.float32(mu) - np.float32(1.0)
    return inv_mulaw(y, mu)


def inv_preemphasis(x, coef=0.85):
    """y[n] = fl(x[n] + fl(coef * y[n-1])) in float32 -- the direct-form-II-transposed loop scipy.signal.lfilter runs
    for b=[1], a=[1,-coef] of dtype float32 (tests compare the two)."""
    x = np.asarray(x, dtype=np.float32)
    y = np.empty_like(x)
    z = np.float32(0.0)
    cf = np.float32(coef)
    for n in range(x.shape[0]):
        y[n] = x[n] + z
        z = cf * y[n]
    return y


def to_int16(x):
    x = np.asarray(x, dtype=np.float32)
    return (x * 32767).astype(np.int16)


def decode(y, lengths=None, input_type="raw", quantize_channels=65536, preemphasis_coef=0.0, global_gain_scale=0.0):
    """y: (B,T) float32 scalars, or (B,T) integer class ids for "mulaw-quantize".
    Returns (float waveforms (B,T), int16 waveforms (B,T)); entries beyond lengths[b] are zero."""
    B, T = y.shape
    mu = quantize_channels - 1
    if input_type == "mulaw-quantize":
        out = np.stack([inv_mulaw_quantize(y[b], mu) for b in range(B)])
    elif input_type == "mulaw":
        out = np.stack([inv_mulaw(y[b], mu) for b in range(B)])
    else:
        out = np.asarray(y, dtype=np.float32).copy()
    if preemphasis_coef != 0.0:
        out = np.stack([inv_preemphasis(out[b], preemphasis_coef) for b in range(B)])
    if global_gain_scale > 0:
        out = (out / np.float32(global_gain_scale)).astype(np.float32)
    pcm = to_int16(np.clip(out, -1.0, 1.0))
    if lengths is not None:
        for b in range(B):
            out[b, lengths[b]:] = 0
            pcm[b, lengths[b]:] = 0
    return out, pcm
