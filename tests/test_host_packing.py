# coding: utf-8
"""Host logic of libwn.so without a GPU: the C ABI loads and exports every declared symbol, the
planner's numbers match SURVEY.md 8(d), and the per-block packed weight image -- read back through
an independent numpy interpreter of the documented layout that replays the kernel's staged
dataflow (row ownership, current-tap / queued older-tap split, skip accumulation, head) --
reproduces the reference's head outputs from the golden vectors."""
import ctypes as C
import math
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, ROOT
from helpers import GoldenCase
from wavenet_vocoder_b200 import _native as N
from wavenet_vocoder_b200.engine import make_config, weights_struct

SMEM = 232448
NSM = 148


def cfg_for(gc, num_ctas=0):
    kw = gc.kw
    return make_config(layers=kw["layers"], stacks=kw["stacks"], residual_channels=kw["residual_channels"],
                       gate_channels=kw["gate_channels"], skip_out_channels=kw["skip_out_channels"],
                       out_channels=kw["out_channels"], kernel_size=kw.get("kernel_size", 3),
                       cin_channels=kw.get("cin_channels", -1), gin_channels=kw.get("gin_channels", -1),
                       scalar_input=kw.get("scalar_input", False),
                       output_distribution=kw.get("output_distribution", "Logistic"), num_ctas=num_ctas)


def plan_of(cfg, batch=1, sms=NSM, smem=SMEM):
    info = N.wn_plan_info()
    N.check(N.lib().wn_plan_only(C.byref(cfg), batch, sms, smem, C.byref(info)))
    return info.as_dict()


def test_library_exports_every_declared_symbol():
    hdr = open(ROOT + "/include/wn.h").read()
    declared = set(re.findall(r"\b(wn_[a-z_]+)\s*\(", hdr))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.wn_abi_version() == N.WN_ABI_VERSION
    # the transcribed ctypes structs have the layout the library was compiled with
    sizes = (C.c_int32 * 5)()
    assert lib.wn_struct_sizes(sizes, 5) == 5
    assert list(sizes) == [C.sizeof(t) for t in (N.wn_config, N.wn_weights, N.wn_generate_args, N.wn_plan_info,
                                                   N.wn_upsampler)]


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = cfg_for(GoldenCase("mol_cond"))
    h = C.c_void_p()
    rc = N.lib().wn_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CUDA device" in N.lib().wn_last_error()
    from wavenet_vocoder_b200 import WaveNet
    m = WaveNet(out_channels=30, layers=2, stacks=1, residual_channels=8, gate_channels=16,
                skip_out_channels=8, scalar_input=True).eval()
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m.incremental_forward(T=4)


def test_plan_numbers_match_survey_table():
    # config 2 = 49 299 456 FLOP/sample, 98.72 MB of fp32 weights per step
    cfg = make_config(layers=24, stacks=4, residual_channels=512, gate_channels=512, skip_out_channels=256,
                      out_channels=30, kernel_size=3, cin_channels=80, gin_channels=-1, scalar_input=True,
                      output_distribution="Logistic")
    p = plan_of(cfg)
    assert p["flops_per_sample"] == 49299456
    assert p["weight_bytes_per_step"] == 4 * 24681246
    assert p["num_ctas"] == 128 and p["rows_y"] == 2 and p["rows_x"] == 4 and p["rows_skip"] == 2
    assert p["exchanges_per_step"] == 27      # one broadcast per layer + skip + two head stages
    assert p["smem_bytes"] <= SMEM
    assert p["resident_blobs"] + p["ring_slots"] >= 3
    # config 1 (1.73 MB) fits entirely in shared memory: nothing streams.  Config 3 (14.67 MB of
    # weights, +36% row-quad padding and the folded M matrices) keeps 17 of 25 blobs resident.
    cfg3 = make_config(layers=24, stacks=4, residual_channels=128, gate_channels=256, skip_out_channels=128,
                       out_channels=2, kernel_size=3, cin_channels=80, gin_channels=16, scalar_input=True,
                       output_distribution="Normal")
    p3 = plan_of(cfg3)
    assert p3["flops_per_sample"] == 7308032 + 2 * 24 * 256 * 0   # Wg.g is folded once per call
    assert p3["resident_blobs"] >= 16 and p3["smem_bytes"] <= SMEM
    cfg1 = make_config(layers=12, stacks=2, residual_channels=64, gate_channels=128, skip_out_channels=64,
                       out_channels=256, kernel_size=3, cin_channels=-1, gin_channels=-1, scalar_input=False,
                       output_distribution="Logistic")
    p1 = plan_of(cfg1)
    assert p1["flops_per_sample"] == 860160 and p1["num_ctas"] == 64
    assert p1["ring_slots"] == 0 and p1["resident_blobs"] == 13 and p1["streamed_bytes_per_step"] == 0
    # config 5: 30 layers / 3 cycles, dilation up to 512
    cfg5 = make_config(layers=30, stacks=3, residual_channels=256, gate_channels=512, skip_out_channels=256,
                       out_channels=30, kernel_size=3, cin_channels=80, gin_channels=-1, scalar_input=True,
                       output_distribution="Logistic")
    p5 = plan_of(cfg5)
    assert p5["flops_per_sample"] == 34061824
    for b in (1, 2, 4, 8):
        assert plan_of(cfg, b)["batch_tile"] == b and plan_of(cfg, b)["smem_bytes"] <= SMEM


def test_planner_rejects_bad_shapes():
    bad = make_config(layers=5, stacks=2, residual_channels=8, gate_channels=16, skip_out_channels=8,
                      out_channels=30, kernel_size=3, cin_channels=-1, gin_channels=-1, scalar_input=True,
                      output_distribution="Logistic")
    info = N.wn_plan_info()
    assert N.lib().wn_plan_only(C.byref(bad), 1, NSM, SMEM, C.byref(info)) == -1
    assert b"multiple of stacks" in N.lib().wn_last_error()
    bad2 = make_config(layers=4, stacks=2, residual_channels=8, gate_channels=16, skip_out_channels=8,
                       out_channels=31, kernel_size=3, cin_channels=-1, gin_channels=-1, scalar_input=True,
                       output_distribution="Logistic")
    assert N.lib().wn_plan_only(C.byref(bad2), 1, NSM, SMEM, C.byref(info)) == -1


# ------------------------------------------------------------------------------------------------
# independent reading of the packed layout (documented in csrc/wn_plan.h / DESIGN.md)
# ------------------------------------------------------------------------------------------------
def part(rows, P, p):
    q, r = divmod(rows, P)
    return p * q + min(p, r), q + (1 if p < r else 0)


def cdiv(a, b):
    return -(-a // b)


def unquad(grp, nq, K):
    """[quad][k][4 rows] -> (4*nq, K)"""
    return grp.reshape(nq, K, 4).transpose(0, 2, 1).reshape(4 * nq, K)


class PackedModel:
    """Reads the packed image back with its own arithmetic for the layout documented in
    csrc/wn_plan.h: first blob [Zx | zb], layer blobs [Zy | Zx | Xo | Td | Sk | zb | xb | sb],
    tail blob [Td | Sk | sb | Ha | Hab | Hb | Hbb]; every matrix group is [quad][k][4 rows]."""

    def __init__(self, gc, P):
        cfg = cfg_for(gc, num_ctas=P)
        self.gc, self.cfg = gc, cfg
        info = plan_of(cfg)
        assert info["num_ctas"] == P
        self.P = P
        c = gc.cfg
        self.L, self.R, self.G2, self.S, self.O = c.layers, c.residual_channels, c.gate_channels // 2, c.skip_out_channels, c.out_channels
        self.kw, self.C = c.kernel_size, max(c.cin_channels, 0)
        L, R, G2, S, O, kw = self.L, self.R, self.G2, self.S, self.O, self.kw
        NYm, NXm, NSm, NAm, NBm = cdiv(G2, P), cdiv(R, P), cdiv(S, P), cdiv(S, P), cdiv(O, P)
        assert (info["rows_y"], info["rows_x"], info["rows_skip"], info["rows_head_a"], info["rows_head_b"]) == \
            (NYm, NXm, NSm, NAm, NBm)
        RA = 2 * NYm
        nqA, nqD = cdiv(RA, 4), cdiv((kw - 1) * RA, 4)
        nqBO, nqBS, nqHA, nqHB = cdiv(NXm, 4), cdiv(NSm, 4), cdiv(NAm, 4), cdiv(NBm, 4)

        def offsets(sizes):
            return [int(v) for v in np.concatenate([[0], np.cumsum(sizes)])]
        fo = offsets([nqA * R * 4, 4 * nqA])
        lo = offsets([nqA * G2 * 4, nqA * R * 4, nqBO * G2 * 4, nqD * R * 4, nqBS * G2 * 4, 4 * nqA, 4 * nqBO, 4 * nqBS])
        to = offsets([nqD * R * 4, nqBS * G2 * 4, 4 * nqBS, nqHA * S * 4, 4 * nqHA, nqHB * S * 4, 4 * nqHB])
        fb, lb, tb = fo[-1], lo[-1], to[-1]
        assert info["layer_blob_bytes"] == 4 * lb and info["head_blob_bytes"] == 4 * tb
        nmain = info["packed_bytes_per_cta"] // 4
        ncond = info["cond_packed_bytes_per_cta"] // 4
        assert nmain == fb + (L - 1) * lb + tb and ncond == L * nqA * self.C * 4
        w, keep = weights_struct(gc.sd, L, self.C, max(c.gin_channels, 0))
        self.blocks = []
        for p in range(P):
            buf = np.zeros(nmain + ncond, dtype=np.float32)
            N.check(N.lib().wn_pack_cta(C.byref(cfg), 1, NSM, SMEM, C.byref(w), p,
                                        buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size))
            blk = dict(y=part(G2, P, p), x=part(R, P, p), s=part(S, P, p), a=part(S, P, p), b=part(O, P, p), stages=[])
            b0 = buf[:fb]
            blk["stages"].append(dict(Zx=unquad(b0[fo[0]:fo[1]], nqA, R), zb=b0[fo[1]:fo[2]]))
            for s_ in range(1, L):
                b = buf[fb + (s_ - 1) * lb: fb + s_ * lb]
                seg = [b[lo[i]:lo[i + 1]] for i in range(8)]
                blk["stages"].append(dict(Zy=unquad(seg[0], nqA, G2), Zx=unquad(seg[1], nqA, R),
                                          Xo=unquad(seg[2], nqBO, G2), Td=unquad(seg[3], nqD, R),
                                          Sk=unquad(seg[4], nqBS, G2), zb=seg[5], xb=seg[6], sb=seg[7]))
            tbuf = buf[fb + (L - 1) * lb: nmain]
            seg = [tbuf[to[i]:to[i + 1]] for i in range(7)]
            blk["tail"] = dict(Td=unquad(seg[0], nqD, R), Sk=unquad(seg[1], nqBS, G2), sb=seg[2],
                               Ha=unquad(seg[3], nqHA, S), Hab=seg[4], Hb=unquad(seg[5], nqHB, S), Hbb=seg[6])
            blk["cond"] = [unquad(buf[nmain + l * nqA * self.C * 4: nmain + (l + 1) * nqA * self.C * 4], nqA, self.C)
                           for l in range(L)] if self.C else None
            self.blocks.append(blk)
        self.RA = RA
        del keep

    def run_teacher_forced(self, b):
        """Replay the kernel's staged dataflow for utterance b; returns (O,T) head outputs.
        Stage s evaluates layer s from (y_{s-1}, x_{s-1}) with conv1x1_out folded into its current
        tap; the older taps' products and the skip rows of layer s-1 are computed one stage late."""
        gc, L, R, G2, S, O, kw, P, RA = self.gc, self.L, self.R, self.G2, self.S, self.O, self.kw, self.P, self.RA
        w = gc.w
        T = gc.T
        dil = gc.cfg.dilations()
        first_w = w["first_w"].numpy()
        first_b = w["first_b"].numpy()
        x_tf = gc.x_tf.numpy()[b]                                  # (C0, T)
        c_up = gc.t("c_up")
        g_vec = gc.t("g_vec")
        gb = None
        if g_vec is not None:
            gb = [lay["g_w"].numpy() @ g_vec[b].numpy() for lay in w["layers"]]     # (G,) per layer
        rings = [[{tap: np.zeros(((kw - 1 - tap) * dil[l], RA), np.float32) for tap in range(kw - 1)}
                  for l in range(L)] for _ in range(P)]
        out = np.zeros((O, T), np.float32)
        rs2 = np.float32(math.sqrt(0.5))

        def gate(p, blk, l, z_dyn, t):
            y0, ny = blk["y"]
            pre = blk["stages"][l]["zb"][:RA].copy()
            if gb is not None:
                for j in range(ny):
                    pre[2 * j] += gb[l][y0 + j]
                    pre[2 * j + 1] += gb[l][G2 + y0 + j]
            if self.C:
                pre += blk["cond"][l][:RA] @ c_up[b, :, t].numpy()
            for tap in range(kw - 1):
                pre += rings[p][l][tap][t % ((kw - 1 - tap) * dil[l])]
            z = z_dyn + pre
            return [(y0 + j, np.tanh(z[2 * j]) / (1.0 + np.exp(-z[2 * j + 1]))) for j in range(ny)]

        def queue_taps(p, Td, layer, xvec, t):
            for tap in range(kw - 1):
                D = (kw - 1 - tap) * dil[layer]
                rings[p][layer][tap][t % D] = Td[tap * RA:(tap + 1) * RA] @ xvec

        for t in range(T):
            x_prev = first_w @ x_tf[:, t] + first_b                # x_0, known to every block
            y_prev = np.zeros(G2, np.float32)
            for p, blk in enumerate(self.blocks):                  # stage 0
                for k, v in gate(p, blk, 0, blk["stages"][0]["Zx"][:RA] @ x_prev, t):
                    y_prev[k] = v
            skipacc = [None] * P
            for s_ in range(1, L):
                y_new = np.zeros(G2, np.float32)
                x_new = np.zeros(R, np.float32)
                for p, blk in enumerate(self.blocks):
                    st = blk["stages"][s_]
                    for k, v in gate(p, blk, s_, st["Zy"][:RA] @ y_prev + st["Zx"][:RA] @ x_prev, t):
                        y_new[k] = v
                    x0, nx = blk["x"]
                    x_new[x0:x0 + nx] = (st["Xo"][:nx] @ y_prev + st["xb"][:nx] + x_prev[x0:x0 + nx]) * rs2
                    queue_taps(p, st["Td"], s_ - 1, x_prev, t)     # deferred
                    s0, ns = blk["s"]
                    h = st["Sk"][:ns] @ y_prev + st["sb"][:ns]
                    skipacc[p] = h if s_ == 1 else skipacc[p] + h
                y_prev, x_prev = y_new, x_new
            sk = np.zeros(S, np.float32)
            for p, blk in enumerate(self.blocks):                  # stage L
                tl = blk["tail"]
                s0, ns = blk["s"]
                h = tl["Sk"][:ns] @ y_prev + tl["sb"][:ns]
                tot = h if L == 1 else skipacc[p] + h
                sk[s0:s0 + ns] = np.maximum(tot * np.float32(math.sqrt(1.0 / L)), 0)
                queue_taps(p, tl["Td"], L - 1, x_prev, t)
            h1 = np.zeros(S, np.float32)
            for blk in self.blocks:
                a0, na = blk["a"]
                h1[a0:a0 + na] = np.maximum(blk["tail"]["Ha"][:na] @ sk + blk["tail"]["Hab"][:na], 0)
            for blk in self.blocks:
                b0, nb = blk["b"]
                out[b0:b0 + nb, t] = blk["tail"]["Hb"][:nb] @ h1 + blk["tail"]["Hbb"][:nb]
        return out


@pytest.mark.parametrize("name,P", [("mol_cond", 5), ("mol_cond", 32), ("mulaw_softmax", 16),
                                    ("gauss_speaker", 3), ("mixgauss", 7), ("mol_upsample", 24)])
def test_packed_image_replays_reference(name, P):
    gc = GoldenCase(name)
    pm = PackedModel(gc, P)
    got = pm.run_teacher_forced(0)
    ref = gc.arr["params_tf"][0]
    assert got.shape == ref.shape
    assert float(np.abs(got - ref).max()) <= 2e-5
