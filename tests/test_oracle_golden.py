# coding: utf-8
"""The oracle against the golden vectors written by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  On the torch build the vectors were made with, the
comparison is bit-exact; a different host CPU may pick other BLAS kernels, so the hard bound is
2e-6 abs on the head outputs (the reference's own online==offline tolerance is 1e-4,
tests/test_model.py:362 in the reference)."""
import pytest
import torch

from conftest import GOLDEN_CASES
from helpers import GoldenCase
from oracle import wavenet_oracle as orc


def test_receptive_field_known_answers():
    # reference tests/test_misc.py:9-13
    assert orc.receptive_field_size(30, 3, 3) == 6139
    assert orc.receptive_field_size(24, 4, 3) == 505
    assert orc.receptive_field_size(12, 2, 3) == 253
    assert orc.receptive_field_size(30, 1, 3, dilation=lambda x: 1) == 61


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_teacher_forced_head_outputs(name):
    gc = GoldenCase(name)
    rec = []
    with torch.no_grad():
        y = orc.incremental_forward(gc.cfg, gc.w, test_inputs=gc.x_tf, c=gc.t("c_up"),
                                    g=gc.t("g_vec"), T=gc.T,
                                    noise=orc.replay_from_predrawn(gc.cfg, gc.noise_tf),
                                    params_out=rec)
    params = torch.stack(rec, dim=-1)
    ref = gc.t("params_tf")
    assert params.shape == ref.shape
    assert float((params - ref).abs().max()) <= 2e-6
    # the reference's own online==offline criterion held when the vectors were made
    assert float(gc.arr["batch_forward_maxdiff"]) < 1e-4
    if gc.cfg.scalar_input:
        assert float((y - gc.t("y_tf")).abs().max()) <= 1e-5
    else:
        agree = (y.argmax(1) == gc.t("y_tf").long()).float().mean().item()
        assert agree >= 0.98


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_free_running_replayed_noise(name):
    gc = GoldenCase(name)
    with torch.no_grad():
        y = orc.incremental_forward(gc.cfg, gc.w, c=gc.t("c_up"), g=gc.t("g_vec"), T=gc.T,
                                    noise=orc.replay_from_predrawn(gc.cfg, gc.noise))
    if gc.cfg.scalar_input:
        ref = gc.t("y_free")
        assert y.shape == ref.shape
        rms = float(((y - ref) ** 2).mean().sqrt())
        assert rms <= 1e-4
    else:
        assert torch.equal(y.argmax(1), gc.t("y_free").long()) or \
            (y.argmax(1) == gc.t("y_free").long()).float().mean().item() >= 0.9


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_predrawn_noise_matches_fixture(name):
    """predraw_noise() reproduces the stored draws (same torch RNG stream) -> seeds are enough."""
    gc = GoldenCase(name)
    n = orc.predraw_noise(gc.cfg, gc.B_free, gc.T, gc.seed)
    for k, v in n.items():
        assert torch.equal(v, gc.noise[k]), k
