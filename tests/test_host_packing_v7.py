# coding: utf-8
"""Host logic of the alternative engine (WN_ENGINE=7, csrc/wn7_*) without a GPU: the C ABI loads and exports every declared symbol, the
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




@pytest.fixture(autouse=True)
def _engine7(monkeypatch):
    monkeypatch.setenv("WN_ENGINE", "7")


def test_plan_of_engine7():
    cfg = make_config(layers=24, stacks=4, residual_channels=512, gate_channels=512, skip_out_channels=256,
                      out_channels=30, kernel_size=3, cin_channels=80, gin_channels=-1, scalar_input=True,
                      output_distribution="Logistic")
    p = plan_of(cfg)
    assert p["engine"] == 7 and p["num_ctas"] == 128 and p["exchanges_per_step"] == 27
    assert p["flops_per_sample"] == 49299456 and p["weight_bytes_per_step"] == 4 * 24681246
    for b in (1, 2, 4, 8):
        q = plan_of(cfg, b)
        assert q["batch_tile"] == b and q["smem_bytes"] <= SMEM


# ------------------------------------------------------------------------------------------------
# independent reading of the packed layout (documented in csrc/wn7_plan.h / DESIGN.md)
# ------------------------------------------------------------------------------------------------
def part(rows, n, p):
    q, r = divmod(rows, n)
    return p * q + min(p, r), q + (1 if p < r else 0)


K_FIRST, K_LAYER, K_TAIL, K_HEAD1, K_HEAD2 = range(5)
J_A0, J_A, J_B, J_D, J_S, J_SL, J_HA, J_HB = range(8)


class PackedModel:
    """Replays the kernel's dataflow from the packed per-block images with its own arithmetic: every pass is two rows
    whose tile [j][row][lane][4 k] is multiplied with the stage vector (k = x_off + 4 (lane + 32 j) + 0..3; vector
    order [y | pad | x] for the layer stages), then finalised exactly as the lanes of the compute warps do (bias,
    conditioning, queued taps, gate, residual, skip accumulation, head)."""

    def __init__(self, gc, P):
        cfg = cfg_for(gc, num_ctas=P)
        self.gc, self.cfg = gc, cfg
        pl, passes = N.plan_passes(cfg, 1, NSM, SMEM)
        assert pl.P == P
        self.pl, self.passes = pl, passes
        info = plan_of(cfg)
        assert info["num_ctas"] == P and info["engine"] == 7 and info["num_passes"] == len(passes)
        c = gc.cfg
        self.L, self.R, self.G2, self.S, self.O = c.layers, c.residual_channels, c.gate_channels // 2, c.skip_out_channels, c.out_channels
        self.kw, self.C = c.kernel_size, max(c.cin_channels, 0)
        cdiv = lambda a, b: -(-a // b)
        assert (info["rows_y"], info["rows_x"], info["rows_skip"]) == (cdiv(self.G2, P), pl.mx, pl.ms)
        assert pl.mx == 2 * cdiv(cdiv(self.R, P), 2) and pl.xoff == 4 * cdiv(self.G2, 4)
        nmain, ncond, nbias = pl.cta_w_floats, pl.cta_cw_floats, pl.cta_b_floats
        assert info["packed_bytes_per_cta"] == 4 * nmain and info["cond_packed_bytes_per_cta"] == 4 * ncond
        assert info["bias_packed_bytes_per_cta"] == 4 * nbias
        assert nmain == pl.fb_floats + (self.L - 1) * pl.lb_floats + pl.tb_floats
        w, keep = weights_struct(gc.sd, self.L, self.C, max(c.gin_channels, 0))
        self.img = []
        for p in range(P):
            buf = np.zeros(nmain + ncond + nbias, dtype=np.float32)
            N.check(N.lib().wn_pack_cta(C.byref(cfg), 1, NSM, SMEM, C.byref(w), p,
                                        buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size))
            self.img.append(dict(w=buf[:nmain], cw=buf[nmain:nmain + ncond], b=buf[nmain + ncond:]))
        del keep

    def blob(self, p, stage):
        pl = self.pl
        i = min(stage, self.L)
        off = 0 if i == 0 else pl.fb_floats + (i - 1) * pl.lb_floats
        n = pl.fb_floats if i == 0 else (pl.lb_floats if i < self.L else pl.tb_floats)
        return self.img[p]["w"][off:off + n]

    def run_stage(self, kind, stage, vec):
        """vec: the stage input vector in xin order.  Returns per block a list of (pass, [sum row 0, sum row 1])."""
        pl = self.pl
        x = np.zeros(pl.xin_vals, np.float32)
        x[:len(vec)] = vec
        res = []
        for p in range(pl.P):
            blob = self.blob(p, stage)
            out = []
            for wv in range(8):
                b0 = pl.pass_begin[kind][wv]
                for ps in self.passes[b0:b0 + pl.pass_count[kind][wv]]:
                    tile = blob[ps.w_off:ps.w_off + ps.nit * 256].reshape(ps.nit, 2, 32, 4)
                    ks = ps.x_off + 4 * (np.arange(32)[None, :, None] + 32 * np.arange(ps.nit)[:, None, None]) + np.arange(4)[None, None, :]
                    sums = np.einsum("jrlk,jlk->r", tile.astype(np.float64), x[ks].astype(np.float64)).astype(np.float32)
                    out.append((ps, sums))
            res.append(out)
        return res

    def run_teacher_forced(self, b):
        gc, pl, L, kw = self.gc, self.pl, self.L, self.kw
        P = pl.P
        G2, R, S, O = self.G2, self.R, self.S, self.O
        my, mx, ms, qA, xoff = pl.my, pl.mx, pl.ms, pl.qA, pl.xoff
        w = gc.w
        T = gc.T
        dil = gc.cfg.dilations()
        first_w = w["first_w"].numpy()
        first_b = w["first_b"].numpy()
        x_tf = gc.x_tf.numpy()[b]
        c_up = gc.t("c_up")
        g_vec = gc.t("g_vec")
        gb = [lay["g_w"].numpy() @ g_vec[b].numpy() for lay in w["layers"]] if g_vec is not None else None
        rs2 = np.float32(math.sqrt(0.5))
        rings = [[{tap: np.zeros(((kw - 1 - tap) * dil[l], 2 * my), np.float32) for tap in range(kw - 1)}
                  for l in range(L)] for _ in range(P)]
        out = np.zeros((O, T), np.float32)
        own = [dict(y=part(G2, P, p), x=part(R, P, p), s=part(S, P, p), a=part(S, P, p), b=part(O, P, p)) for p in range(P)]

        def pre_of(p, l, t):
            bias = self.img[p]["b"]
            y0, ny = own[p]["y"]
            pre = bias[pl.bo_zb + l * 2 * my: pl.bo_zb + (l + 1) * 2 * my].copy()
            for j in range(ny):
                if gb is not None:
                    pre[2 * j] += gb[l][y0 + j]
                    pre[2 * j + 1] += gb[l][G2 + y0 + j]
            if self.C:
                cw = self.img[p]["cw"][l * qA * self.C * 4:(l + 1) * qA * self.C * 4].reshape(qA, self.C, 4)
                cw = cw.transpose(0, 2, 1).reshape(4 * qA, self.C)[:2 * my]
                pre += cw @ c_up[b, :, t].numpy()
            for tap in range(kw - 1):
                pre += rings[p][l][tap][t % ((kw - 1 - tap) * dil[l])]
            return pre

        for t in range(T):
            x_prev = (first_w @ x_tf[:, t] + first_b).astype(np.float32)
            y_prev = np.zeros(G2, np.float32)
            skipacc = np.zeros((P, ms), np.float32)
            for s_ in range(0, L + 1):
                kind = K_FIRST if s_ == 0 else (K_LAYER if s_ < L else K_TAIL)
                vec = np.zeros(xoff + R, np.float32)
                vec[:G2] = y_prev
                vec[xoff:xoff + R] = x_prev
                res = self.run_stage(kind, s_, vec)
                y_new, x_new = np.zeros(G2, np.float32), x_prev.copy()
                sk = np.zeros(S, np.float32)
                for p in range(P):
                    bias = self.img[p]["b"]
                    y0, ny = own[p]["y"]
                    x0, nx = own[p]["x"]
                    s0, ns = own[p]["s"]
                    pre = pre_of(p, s_, t) if s_ < L else None
                    for ps, v in res[p]:
                        if ps.job in (J_A0, J_A):
                            i = ps.idx
                            if i < ny:
                                za, zb = v[0] + pre[2 * i], v[1] + pre[2 * i + 1]
                                y_new[y0 + i] = np.tanh(za) / (1.0 + np.exp(-zb))
                        elif ps.job == J_B:
                            for r in range(2):
                                j = ps.idx + r
                                if j < nx:
                                    x_new[x0 + j] = (v[r] + bias[pl.bo_xb + s_ * mx + j] + x_prev[x0 + j]) * rs2
                        elif ps.job == J_D:
                            tap, i = divmod(ps.idx, my)
                            D = (kw - 1 - tap) * dil[s_ - 1]
                            rings[p][s_ - 1][tap][t % D][2 * i:2 * i + 2] = v
                        elif ps.job == J_S:
                            for r in range(2):
                                j = ps.idx + r
                                if j < ns:
                                    h = v[r] + bias[pl.bo_sb + (s_ - 1) * ms + j]
                                    skipacc[p, j] = h if s_ == 1 else skipacc[p, j] + h
                        elif ps.job == J_SL:
                            for r in range(2):
                                j = ps.idx + r
                                if j < ns:
                                    tot = v[r] + bias[pl.bo_sb + (L - 1) * ms + j]
                                    if L >= 2:
                                        tot = skipacc[p, j] + tot
                                    sk[s0 + j] = max(tot * np.float32(math.sqrt(1.0 / L)), 0)
                        else:
                            raise AssertionError(ps.job)
                if s_ < L:
                    y_prev = y_new
                    if s_ >= 1:
                        x_prev = x_new
            h1 = np.zeros(S, np.float32)
            for p, lst in enumerate(self.run_stage(K_HEAD1, L + 1, sk)):
                a0, na = own[p]["a"]
                for ps, v in lst:
                    assert ps.job == J_HA
                    for r in range(2):
                        j = ps.idx + r
                        if j < na:
                            h1[a0 + j] = max(v[r] + self.img[p]["b"][pl.bo_ha + j], 0)
            for p, lst in enumerate(self.run_stage(K_HEAD2, L + 2, h1)):
                b0, nb = own[p]["b"]
                for ps, v in lst:
                    assert ps.job == J_HB
                    for r in range(2):
                        j = ps.idx + r
                        if j < nb:
                            out[b0 + j, t] = v[r] + self.img[p]["b"][pl.bo_hb + j]
        return out


@pytest.mark.parametrize("name,P", [("mol_cond", 5), ("mol_cond", 16), ("mulaw_softmax", 16),
                                    ("gauss_speaker", 3), ("mixgauss", 7), ("mol_upsample", 12)])
def test_packed_image_replays_reference(name, P):
    gc = GoldenCase(name)
    pm = PackedModel(gc, P)
    got = pm.run_teacher_forced(0)
    ref = gc.arr["params_tf"][0]
    assert got.shape == ref.shape
    assert float(np.abs(got - ref).max()) <= 2e-5


def test_pass_lists_cover_every_row_once():
    """Every row pair of every job appears in exactly one pass per stage kind, critical passes precede deferred ones
    in every warp, and tiles do not overlap inside a blob."""
    cfg = make_config(layers=24, stacks=4, residual_channels=512, gate_channels=512, skip_out_channels=256,
                      out_channels=30, kernel_size=3, cin_channels=80, gin_channels=-1, scalar_input=True,
                      output_distribution="Logistic")
    for batch in (1, 8):
        pl, passes = N.plan_passes(cfg, batch)
        assert (pl.P, pl.BT, pl.my, pl.mx, pl.ms, pl.mo) == (128, batch, 2, 4, 2, 2)
        want = {K_FIRST: {J_A0: [0, 1]}, K_LAYER: {J_A: [0, 1], J_B: [0, 2], J_D: [0, 1, 2, 3], J_S: [0]},
                K_TAIL: {J_SL: [0], J_D: [0, 1, 2, 3]}, K_HEAD1: {J_HA: [0]}, K_HEAD2: {J_HB: [0]}}
        for kind in range(5):
            seen = {}
            spans = []
            for wv in range(8):
                b0, n, nc = pl.pass_begin[kind][wv], pl.pass_count[kind][wv], pl.pass_crit[kind][wv]
                for i, ps in enumerate(passes[b0:b0 + n]):
                    assert (ps.deferred == 0) == (i < nc)
                    assert ps.x_off % 4 == 0 and ps.x_off + 128 * ps.nit <= pl.xin_vals
                    spans.append((ps.w_off, ps.w_off + ps.nit * 256))
                    seen.setdefault(ps.job, []).append(ps.idx)
            assert {j: sorted(v) for j, v in seen.items()} == want[kind]
            if kind in (K_FIRST, K_LAYER):
                spans.sort()
                assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
