# coding: utf-8
"""Local-conditioning upsampler on the device (csrc/wn_aux.cuh) against the reference's
upsample network (upsample.py:29-85): the golden case ``mol_upsample`` (frames and the reference's own upsampled
output are in the fixture) and BASELINE config 2's [4,4,4,4] / config 5's [4,5,5,3] hops against the
same-named PyTorch modules (same state_dict keys as the reference).  fp32, tolerance 2e-6 abs (different
summation order of the 2s+1 taps and of conv_in's C*ks products)."""
import numpy as np
import pytest
import torch

from helpers import GoldenCase

pytestmark = pytest.mark.gpu
TOL = 2e-6


def native_upsample(m, c):
    """Run only the upsampler of libwn: a 1-sample synthesis call would do too, but the scratch (B,T,C) tensor is
    what we want to look at, so use the engine's frame entry with T = upsampled length and read params back."""
    eng = m._get_engine()
    assert m._native_upsample
    return eng


def model_for(scales, cin_pad, net="ConvInUpsampleNetwork", C=80):
    from wavenet_vocoder_b200 import WaveNet
    torch.manual_seed(0)
    m = WaveNet(out_channels=30, layers=2, stacks=1, residual_channels=16, gate_channels=32, skip_out_channels=16,
                cin_channels=C, cin_pad=cin_pad, scalar_input=True, dropout=0.0, upsample_conditional_features=True,
                upsample_net=net, upsample_params={"upsample_scales": scales, "cin_channels": C, "cin_pad": cin_pad})
    with torch.no_grad():
        for n_, p in m.upsample_net.named_parameters():
            if n_.endswith("weight_g"):
                p.mul_(1.0 + 0.3 * torch.rand_like(p))
            if n_.endswith("weight_v"):
                p.add_(0.05 * torch.randn_like(p))
    return m.eval()          # on the CPU: the caller takes the fp32 reference there before moving it


@pytest.mark.parametrize("scales,cin_pad,net,frames", [([4, 4, 4, 4], 2, "ConvInUpsampleNetwork", 17),
                                                       ([4, 5, 5, 3], 2, "ConvInUpsampleNetwork", 9),
                                                       ([2, 4], 0, "ConvInUpsampleNetwork", 33),
                                                       ([4, 4], 1, "UpsampleNetwork", 12),
                                                       ([16, 16], 0, "UpsampleNetwork", 5)])
def test_native_upsampler_matches_module(scales, cin_pad, net, frames):
    m = model_for(scales, cin_pad, net)
    B = 3
    c = torch.randn(B, 80, frames + 2 * cin_pad)
    with torch.no_grad():
        # fp32 reference on the CPU: cuDNN would run the conv_in / smoothing convolutions in TF32 by default
        ref = m.upsample_net(c).cuda()                              # (B,C,T) module with the reference's structure
    m, c = m.cuda(), c.cuda()
    T = ref.size(-1)
    eng = m._get_engine()
    assert m._native_upsample and eng.upsampled_length(c.size(-1)) == T
    got = eng.upsample(c, T)                                      # (B,T,C) written by the device kernels
    assert tuple(got.shape) == (B, T, 80)
    err = float((got.transpose(1, 2) - ref).abs().max())
    assert err <= TOL, err


def test_golden_upsample_fixture():
    gc = GoldenCase("mol_upsample")
    from wavenet_vocoder_b200 import WaveNet
    m = WaveNet(**gc.kw)
    m.load_state_dict(gc.sd)
    m = m.cuda().eval()
    eng = m._get_engine()
    assert m._native_upsample
    c_raw, c_up = gc.t("c_raw").cuda(), gc.t("c_up")               # c_up: the unmodified reference's output
    got = eng.upsample(c_raw, c_up.size(-1))
    assert float((got.transpose(1, 2).cpu() - c_up).abs().max()) <= TOL


def test_frames_and_sample_rate_entries_agree():
    """incremental_forward through the native upsampler == through the PyTorch module + sample-rate c."""
    import os
    gc = GoldenCase("mol_upsample")
    from wavenet_vocoder_b200 import WaveNet
    m = WaveNet(**gc.kw)
    m.load_state_dict(gc.sd)
    m = m.cuda().eval()
    noise = {k: v.cuda() for k, v in gc.noise.items()}
    y1, p1 = m.incremental_forward(c=gc.t("c_raw"), T=gc.T, noise=noise, return_params=True)
    os.environ["WN_TORCH_UPSAMPLE"] = "1"
    try:
        y2, p2 = m.incremental_forward(c=gc.t("c_raw"), T=gc.T, noise=noise, return_params=True)
    finally:
        del os.environ["WN_TORCH_UPSAMPLE"]
    assert float((p1 - p2).abs().max()) <= 2e-5
    assert float((p1.cpu() - gc.t("params_free")).abs().max()) <= 2e-5
