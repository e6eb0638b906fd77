# coding: utf-8
"""Directory-level synthesis: the part of the reference's ``evaluate.py`` that sits either side of
the hot path -- read ``*-feats.npy`` mel files, batch them, synthesise,
decode and write 16-bit wav files -- re-organised for the engine:

* the reference pads every utterance of a DataLoader batch to the longest one and loops batches on one
  device (evaluate.py:50-58, :162-204).  Here utterances are sharded over ranks by total length
  (``parallel.shard_utterances``) and, inside a rank, sorted by length and grouped into launches of at
  most ``tile`` utterances, so the padding inside a launch is small;
* decode follows synthesis.py:66-84 / evaluate.py:215-251: class ids -> inverse mu-law, optional inverse
  pre-emphasis, gain, trim to the utterance's own length, clip, int16.

File formats are the reference's: ``<name>-feats.npy`` = (frames, num_mels) float32 written by
``datasets/wavallin.py`` (np.save, :104-107); output ``<name>_gen.wav`` (evaluate.py:229-231).
"""
from __future__ import annotations

import os
from glob import glob
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .parallel import shard_utterances, tile_batches


def inv_mulaw(y: np.ndarray, mu: int = 255) -> np.ndarray:
    """[-1,1] mu-law companded -> linear (what nnmnkwii's ``inv_mulaw`` does for synthesis.py:71-74)."""
    y = np.asarray(y, dtype=np.float32)
    return np.sign(y) * (1.0 / mu) * ((1.0 + mu) ** np.abs(y) - 1.0)


def inv_mulaw_quantize(idx: np.ndarray, mu: int = 255) -> np.ndarray:
    """class ids in [0, mu] -> linear waveform in [-1, 1] (synthesis.py:66-70)."""
    y = 2.0 * np.asarray(idx, dtype=np.float32) / mu - 1.0
    return inv_mulaw(y, mu)


def inv_preemphasis(x: np.ndarray, coef: float = 0.85) -> np.ndarray:
    """audio.py:57-58: y[n] = x[n] + coef * y[n-1]."""
    from scipy import signal
    return signal.lfilter([1.0], [1.0, -coef], x).astype(np.float32)


def to_int16(x: np.ndarray) -> np.ndarray:
    """evaluate.py:43-48."""
    if x.dtype == np.int16:
        return x
    x = np.asarray(x, dtype=np.float32)
    assert x.min() >= -1 and x.max() <= 1.0
    return (x * 32767).astype(np.int16)


def list_feature_files(data_dir: str) -> List[str]:
    files = sorted(glob(os.path.join(data_dir, "*-feats.npy")))
    if not files:
        raise FileNotFoundError("no *-feats.npy under %s" % data_dir)
    return files


def collate(feats: Sequence[np.ndarray], cin_pad: int) -> torch.Tensor:
    """(frames_i, D) arrays -> (B, D, max_frames + 2*cin_pad): zero padded to the longest (evaluate.py:50-58)
    then edge-replicated by cin_pad on both sides (evaluate.py:163-164)."""
    max_len = max(f.shape[0] for f in feats)
    D = feats[0].shape[1]
    c = np.zeros((len(feats), max_len, D), dtype=np.float32)
    for i, f in enumerate(feats):
        c[i, :f.shape[0]] = f
    ct = torch.from_numpy(c).transpose(1, 2).contiguous()
    if cin_pad > 0:
        ct = F.pad(ct, (cin_pad, cin_pad), mode="replicate")
    return ct


def decode(y_hat: torch.Tensor, input_type: str = "raw", quantize_channels: int = 65536,
           postprocess: Optional[str] = None, global_gain_scale: float = 0.0) -> np.ndarray:
    """Model output (B,C,T) -> float waveforms (B,T) as synthesis.py:66-84 does."""
    B = y_hat.size(0)
    if input_type == "mulaw-quantize":
        out = inv_mulaw_quantize(y_hat.max(1)[1].view(B, -1).cpu().numpy(), quantize_channels - 1)
    elif input_type == "mulaw":
        out = inv_mulaw(y_hat.view(B, -1).cpu().numpy(), quantize_channels - 1)
    else:
        out = y_hat.view(B, -1).cpu().numpy().astype(np.float32)
    if postprocess == "inv_preemphasis":
        out = np.stack([inv_preemphasis(o) for o in out])
    if global_gain_scale > 0:
        out = out / global_gain_scale
    return out


def synthesize_directory(model, data_dir: str, dst_dir: str, *, hop_size: int, cin_pad: int = 0,
                         sample_rate: int = 22050, tile: int = 4, rank: int = 0, world: int = 1,
                         input_type: str = "raw", quantize_channels: int = 65536,
                         postprocess: Optional[str] = None, global_gain_scale: float = 0.0,
                         synth: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None,
                         write: bool = True) -> Dict[str, np.ndarray]:
    """Synthesise this rank's share of ``data_dir`` and write ``<name>_gen.wav`` into ``dst_dir``.
    ``synth(c, T)`` defaults to ``model.incremental_forward(c=c, T=T)`` (c: (B, D, frames + 2*cin_pad)).
    Returns {name: int16 waveform} for the utterances handled by this rank."""
    from scipy.io import wavfile
    files = list_feature_files(data_dir)
    feats = [np.load(f).astype(np.float32) for f in files]
    lengths = [f.shape[0] * hop_size for f in feats]
    mine = shard_utterances(lengths, world)[rank]
    if synth is None:
        def synth(c, T):
            with torch.no_grad():
                return model.incremental_forward(c=c, T=T)
    if write:
        os.makedirs(dst_dir, exist_ok=True)
    results: Dict[str, np.ndarray] = {}
    for launch in tile_batches(mine, lengths, tile):
        c = collate([feats[i] for i in launch], cin_pad)
        T = (c.shape[-1] - 2 * cin_pad) * hop_size
        waves = decode(synth(c, T), input_type, quantize_channels, postprocess, global_gain_scale)
        for row, i in enumerate(launch):
            gen = np.clip(waves[row][:lengths[i]], -1.0, 1.0)          # trim the batch padding (evaluate.py:215,247)
            name = os.path.splitext(os.path.basename(files[i]))[0].replace("-feats", "")
            pcm = to_int16(gen.astype(np.float32))
            results[name] = pcm
            if write:
                wavfile.write(os.path.join(dst_dir, "%s_gen.wav" % name), sample_rate, pcm)
    return results
