# coding: utf-8
"""The CPU arm of bench.py times the reference's own package when it has been staged under the git-ignored
oracle/_ref/ (by __graft_entry__.build(), which has /root/reference in the build container; the directory
travels to the GPU box with the snapshot).  These tests pin the oracle port against that staged package
wherever it is present -- same weights, same conditioning, same torch seed -> torch.equal -- so the
"port" and "reference" kinds of cpu_baseline are interchangeable, and they re-check the golden vectors
against it.  One copy runs in the CPU suite, one is marked gpu so it also runs on the GPU box."""
import os
import sys
import warnings

import pytest
import torch

from conftest import GOLDEN_CASES, ROOT
from helpers import GoldenCase
from oracle import wavenet_oracle as orc

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = os.path.isfile(os.path.join(REF_DIR, "wavenet_vocoder", "wavenet.py"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not staged (run __graft_entry__.build() where "
                                                    "/root/reference exists)")


def ref_model(gc):
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import wavenet_vocoder
        m = wavenet_vocoder.WaveNet(**gc.kw).eval()
        m.load_state_dict(gc.sd)
    return m


def check_case(name):
    gc = GoldenCase(name)
    m = ref_model(gc)
    cfg, w = gc.cfg, gc.w
    c_raw, c_up, g_ids = gc.t("c_raw"), gc.t("c_up"), gc.t("g_ids")
    g_vec = orc.embed_speaker(w, g_ids) if g_ids is not None else None
    T = min(gc.T, 48)
    if c_raw is not None and gc.kw.get("upsample_conditional_features"):
        T = gc.T                                  # the upsample network fixes the length
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(11)
        y_ref = m.incremental_forward(c=c_raw if c_raw is None or c_raw.size(-1) != gc.T or T == gc.T else c_raw[..., :T],
                                      g=g_ids, T=T)
        torch.manual_seed(11)
        y_orc = orc.incremental_forward(cfg, w, c=None if c_up is None else c_up[..., :T], g=g_vec, T=T)
    assert torch.equal(y_ref, y_orc), name
    # and the committed golden vector (full length, seed of make_golden.py) is what the staged package produces
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(gc.seed)
        y_full = m.incremental_forward(c=c_raw, g=g_ids, T=gc.T)
    ref = gc.t("y_free")
    if cfg.scalar_input:
        assert torch.equal(y_full, ref), name
    else:
        assert torch.equal(y_full.argmax(1), ref.long()), name


@needs_ref
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_equals_staged_reference(name):
    check_case(name)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mol_cond", "gauss_speaker"])
def test_port_equals_staged_reference_on_gpu_box(name):
    check_case(name)


@needs_ref
def test_bench_reference_arm_uses_the_staged_package():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.load_reference() is not None
    m = bench.build_model()
    v, dt, n, kind = bench.time_cpu(m, 60, 5, 2, budget_s=2.0)
    assert kind == "reference" and v > 0 and n >= 50
