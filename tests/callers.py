# coding: utf-8
"""The reference's CALLERS of the path, restated for the drop-in tests: ``train.sanity_check``
(train.py:72-87), ``synthesis.batch_wavegen`` (synthesis.py:42-86) and ``synthesis.wavegen`` (synthesis.py:101-189).
The originals import docopt / nnmnkwii / librosa / tensorboardX, which are not installed here, so the bodies are
restated with those dependencies replaced by the small stand-ins below (``HP`` for hparams, ``audio`` and ``P`` backed
by oracle/decode_oracle.py).  Everything they do with the MODEL is kept call for call: ``sanity_check`` ->
``eval()`` -> ``make_generation_fast_()`` -> ``.to(device)`` of the inputs -> ``incremental_forward(c=, g=, T=, tqdm=,
softmax=True, quantize=True, log_scale_min=)`` -> post-processing of the returned tensor.  TEST INFRASTRUCTURE."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import decode_oracle as dorc


def make_hparams(**kw):
    hp = dict(input_type="raw", quantize_channels=65536, upsample_conditional_features=False, cin_pad=0,
              hop_size=256, log_scale_min=-16.0, postprocess="", global_gain_scale=1.0, cin_channels=80)
    hp.update(kw)
    return SimpleNamespace(**hp)


def is_mulaw_quantize(s):
    return s == "mulaw-quantize"


def is_mulaw(s):
    return s == "mulaw"


class P:                      # nnmnkwii.preprocessing stand-in (see oracle/decode_oracle.py)
    inv_mulaw = staticmethod(dorc.inv_mulaw)
    inv_mulaw_quantize = staticmethod(dorc.inv_mulaw_quantize)

    @staticmethod
    def mulaw_quantize(x, mu=255):
        y = np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)
        return int((y + 1) / 2 * mu)


class audio:                  # audio.py stand-in
    hp = None

    @staticmethod
    def get_hop_size():
        return audio.hp.hop_size

    @staticmethod
    def inv_preemphasis(x, coef=0.85):
        return dorc.inv_preemphasis(x, coef)


def to_categorical(y, num_classes):
    out = np.zeros((1, num_classes), dtype=np.float32)
    out[0, int(y)] = 1.0
    return out


def sanity_check(model, c, g):                                   # train.py:72-87
    if model.has_speaker_embedding():
        if g is None:
            raise RuntimeError("WaveNet expects speaker embedding, but speaker-id is not provided")
    else:
        if g is not None:
            raise RuntimeError("WaveNet expects no speaker embedding, but speaker-id is provided")
    if model.local_conditioning_enabled():
        if c is None:
            raise RuntimeError("WaveNet expects conditional features, but not given")
    else:
        if c is not None:
            raise RuntimeError("WaveNet expects no conditional features, but given")


def batch_wavegen(model, hparams, device, c=None, g=None, fast=True, tqdm=lambda x: x):      # synthesis.py:42-86
    audio.hp = hparams
    sanity_check(model, c, g)
    assert c is not None
    B = c.shape[0]
    model.eval()
    if fast:
        model.make_generation_fast_()
    g = None if g is None else g.to(device)
    c = None if c is None else c.to(device)
    if hparams.upsample_conditional_features:
        length = (c.shape[-1] - hparams.cin_pad * 2) * audio.get_hop_size()
    else:
        length = c.shape[-1]
    with torch.no_grad():
        y_hat = model.incremental_forward(c=c, g=g, T=length, tqdm=tqdm, softmax=True, quantize=True,
                                          log_scale_min=hparams.log_scale_min)
    if is_mulaw_quantize(hparams.input_type):
        y_hat = y_hat.max(1)[1].view(B, -1).float().cpu().data.numpy()
        for i in range(B):
            y_hat[i] = P.inv_mulaw_quantize(y_hat[i], hparams.quantize_channels - 1)
    elif is_mulaw(hparams.input_type):
        y_hat = y_hat.view(B, -1).cpu().data.numpy()
        for i in range(B):
            y_hat[i] = P.inv_mulaw(y_hat[i], hparams.quantize_channels - 1)
    else:
        y_hat = y_hat.view(B, -1).cpu().data.numpy()
    if hparams.postprocess is not None and hparams.postprocess not in ["", "none"]:
        for i in range(B):
            y_hat[i] = getattr(audio, hparams.postprocess)(y_hat[i])
    if hparams.global_gain_scale > 0:
        for i in range(B):
            y_hat[i] /= hparams.global_gain_scale
    return y_hat


def wavegen(model, hparams, device, length=None, c=None, g=None, initial_value=None, fast=False,
            tqdm=lambda x: x):                                                               # synthesis.py:101-189
    audio.hp = hparams
    sanity_check(model, c, g)
    model.eval()
    if fast:
        model.make_generation_fast_()
    if c is None:
        assert length is not None
    else:
        if c.ndim != 2:
            raise RuntimeError("Expected 2-dim shape (T, {}) for the conditional feature".format(hparams.cin_channels))
        Tc = c.shape[0]
        upsample_factor = audio.get_hop_size()
        length = Tc * upsample_factor
        if not hparams.upsample_conditional_features:
            c = np.repeat(c, upsample_factor, axis=0)
        c = torch.FloatTensor(c.T).unsqueeze(0)
    if initial_value is None:
        if is_mulaw_quantize(hparams.input_type):
            initial_value = P.mulaw_quantize(0, hparams.quantize_channels - 1)
        else:
            initial_value = 0.0
    if is_mulaw_quantize(hparams.input_type):
        assert initial_value >= 0 and initial_value < hparams.quantize_channels
        initial_input = to_categorical(initial_value, num_classes=hparams.quantize_channels).astype(np.float32)
        initial_input = torch.from_numpy(initial_input).view(1, 1, hparams.quantize_channels)
    else:
        initial_input = torch.zeros(1, 1, 1).fill_(initial_value)
    g = None if g is None else torch.LongTensor([g])
    initial_input = initial_input.to(device)
    g = None if g is None else g.to(device)
    c = None if c is None else c.to(device)
    with torch.no_grad():
        y_hat = model.incremental_forward(initial_input, c=c, g=g, T=length, tqdm=tqdm, softmax=True, quantize=True,
                                          log_scale_min=hparams.log_scale_min)
    if is_mulaw_quantize(hparams.input_type):
        y_hat = y_hat.max(1)[1].view(-1).long().cpu().data.numpy()
        y_hat = P.inv_mulaw_quantize(y_hat, hparams.quantize_channels)
    elif is_mulaw(hparams.input_type):
        y_hat = P.inv_mulaw(y_hat.view(-1).cpu().data.numpy(), hparams.quantize_channels)
    else:
        y_hat = y_hat.view(-1).cpu().data.numpy()
    if hparams.postprocess is not None and hparams.postprocess not in ["", "none"]:
        y_hat = getattr(audio, hparams.postprocess)(y_hat)
    if hparams.global_gain_scale > 0:
        y_hat /= hparams.global_gain_scale
    return y_hat
