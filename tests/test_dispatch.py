# coding: utf-8
"""Host logic around the hot path: feature files in, 16-bit wav out, with
length-bucketed launches and rank sharding.  CPU tests use a stand-in synthesiser; the GPU test
drives the real engine."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as dorc
from wavenet_vocoder_b200 import dispatch as D


def cpu_decode(y_hat, lens):
    """The oracle's decode as the stand-in for the device kernel in the host-logic tests."""
    return dorc.decode(y_hat.reshape(y_hat.size(0), -1).numpy(), lens)[1]


def make_dir(tmp_path, frames, mels=8, seed=0):
    rng = np.random.RandomState(seed)
    for i, n in enumerate(frames):
        np.save(os.path.join(tmp_path, "utt%02d-feats.npy" % i), rng.randn(n, mels).astype(np.float32))
    return str(tmp_path)


def fake_synth(hop):
    def synth(c, T):
        # (B, D, frames+2pad) -> (B,1,T): every sample is tanh of its frame's first mel bin
        B = c.shape[0]
        pad = (c.shape[-1] * hop - T) // (2 * hop)
        core = c[:, 0, pad:c.shape[-1] - pad]
        return torch.tanh(core).repeat_interleave(hop, dim=1).view(B, 1, T)
    return synth


def test_directory_synthesis_trims_and_names(tmp_path):
    frames = [5, 9, 3, 9, 7, 2]
    hop, pad = 4, 2
    src = make_dir(tmp_path, frames)
    dst = os.path.join(src, "out")
    res = D.synthesize_directory(None, src, dst, hop_size=hop, cin_pad=pad, sample_rate=8000, tile=4,
                                 synth=fake_synth(hop), decode=cpu_decode)
    assert sorted(res) == ["utt%02d" % i for i in range(len(frames))]
    from scipy.io import wavfile
    for i, n in enumerate(frames):
        sr, pcm = wavfile.read(os.path.join(dst, "utt%02d_gen.wav" % i))
        assert sr == 8000 and pcm.dtype == np.int16 and len(pcm) == n * hop        # own length, not the tile's
        feats = np.load(os.path.join(src, "utt%02d-feats.npy" % i))
        want = dorc.to_int16(np.tanh(feats[:, 0]).repeat(hop).astype(np.float32))
        assert np.array_equal(pcm, want)


def test_rank_shards_cover_everything_once(tmp_path):
    frames = [30, 4, 17, 17, 9, 25, 6, 12]
    src = make_dir(tmp_path, frames, seed=3)
    parts = [D.synthesize_directory(None, src, src, hop_size=2, tile=3, rank=r, world=3, synth=fake_synth(2),
                                    decode=cpu_decode, write=False) for r in range(3)]
    names = sorted(n for p in parts for n in p)
    assert names == ["utt%02d" % i for i in range(len(frames))]
    loads = [sum(len(v) for v in p.values()) for p in parts]
    assert max(loads) - min(loads) <= 2 * max(frames)


@pytest.mark.gpu
def test_directory_synthesis_on_engine(tmp_path):
    from wavenet_vocoder_b200 import WaveNet
    hop, pad, mels = 8, 1, 8
    torch.manual_seed(0)
    m = WaveNet(out_channels=30, layers=4, stacks=2, residual_channels=16, gate_channels=32, skip_out_channels=16,
                cin_channels=mels, cin_pad=pad, scalar_input=True, dropout=0.0, upsample_conditional_features=True,
                upsample_params={"upsample_scales": [2, 4], "cin_channels": mels, "cin_pad": pad}).cuda().eval()
    src = make_dir(tmp_path, [6, 3, 5, 4, 6], mels=mels)
    torch.manual_seed(1)
    res = D.synthesize_directory(m, src, os.path.join(src, "out"), hop_size=hop, cin_pad=pad, tile=4)
    assert sorted(len(v) for v in res.values()) == sorted(n * hop for n in [6, 3, 5, 4, 6])
    assert all(v.dtype == np.int16 for v in res.values())
    # numerically: the same launches through the class surface + the oracle's decode give the same int16 files
    from wavenet_vocoder_b200.parallel import shard_utterances, tile_batches
    files = D.list_feature_files(src)
    feats = [np.load(f) for f in files]
    lengths = [f.shape[0] * hop for f in feats]
    torch.manual_seed(1)
    for launch in tile_batches(shard_utterances(lengths, 1)[0], lengths, 4):
        c = D.collate([feats[i] for i in launch], pad)
        T = (c.shape[-1] - 2 * pad) * hop
        with torch.no_grad():
            y = m.incremental_forward(c=c, T=T)
        _, pcm = dorc.decode(y.view(len(launch), -1).cpu().numpy(), [lengths[i] for i in launch])
        for row, i in enumerate(launch):
            assert np.array_equal(res["utt%02d" % i], pcm[row][:lengths[i]])
