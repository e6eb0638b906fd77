# coding: utf-8
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["mulaw_softmax", "mol_cond", "mol_upsample", "gauss_speaker", "mixgauss"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the oracle is a chain of tiny CPU GEMVs: with one thread per core of a large host every step pays the fork/join
    # of ~200 threads (measured 0.1-0.2 s per oracle step on the GPU box against 5 ms with 4 threads)
    import torch
    torch.set_num_threads(max(1, min(4, os.cpu_count() or 1)))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
