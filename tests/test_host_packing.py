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
# independent reading of the packed layout (documented in csrc/wn6_plan.h / DESIGN.md)
# ------------------------------------------------------------------------------------------------
def part(rows, n, p):
    q, r = divmod(rows, n)
    return p * q + min(p, r), q + (1 if p < r else 0)


def own(rows, NC, CS, c, r):
    cb, cc = part(rows, NC, c)
    ob, oc = part(cc, CS, r)
    return cb + ob, oc


K_FIRST, K_LAYER, K_TAIL, K_HEAD1, K_HEAD2 = range(5)


class PackedModel:
    """Replays the cluster engine's dataflow from the packed per-block images with its own arithmetic:
    K-slices gathered from the values the rank-r blocks of all clusters published, pass tiles
    [j][lane][4 rows] multiplied with the slice (k = x_off + sub + 16 j), partial sums summed by the row
    owners in rank order, then bias / conditioning / queued taps / gate / residual exactly as the finaliser
    warps do."""

    def __init__(self, gc, P, cluster=0):
        cfg = cfg_for(gc, num_ctas=P)
        cfg.cluster_size = cluster
        self.gc, self.cfg = gc, cfg
        pl, passes = N.plan_passes(cfg, 1, NSM, SMEM)
        assert pl.P == P and pl.NC * pl.CS == P
        self.pl, self.passes = pl, passes
        info = plan_of(cfg)
        assert info["num_ctas"] == P and info["num_clusters"] == pl.NC and info["cluster_size"] == pl.CS
        assert info["engine"] == 6 and info["num_passes"] == len(passes)
        c = gc.cfg
        self.L, self.R, self.G2, self.S, self.O = c.layers, c.residual_channels, c.gate_channels // 2, c.skip_out_channels, c.out_channels
        self.kw, self.C = c.kernel_size, max(c.cin_channels, 0)
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

    def run_stage(self, kind, stage, xin):
        """xin[r]: the K-slice of rank r (values).  Returns the partial sums [P][row][src rank] sent to the three
        finaliser warps of every owner: (F0: gate / skip / head rows, DF: deferred rows, F1: residual rows)."""
        pl = self.pl
        crit = np.zeros((pl.P, pl.nrow_c, pl.CS), np.float32)
        defer = np.zeros((pl.P, pl.nrow_d, pl.CS), np.float32)
        resid = np.zeros((pl.P, max(pl.nrow_x, 1), pl.CS), np.float32)
        for p in range(pl.P):
            c, r = divmod(p, pl.CS)
            blob = self.blob(p, stage)
            x = np.zeros(pl.xin_vals + 64, np.float32)
            x[:len(xin[r])] = xin[r]
            for wv in range(8):
                b0 = pl.pass_begin[kind][wv]
                for ps in self.passes[b0:b0 + pl.pass_count[kind][wv]]:
                    tile = blob[ps.w_off:ps.w_off + ps.nit * 128].reshape(ps.nit, 32, 4)
                    for g in range(2):
                        if ps.owner[g] < 0:
                            continue
                        ks = ps.x_off + np.arange(16)[None, :] + 16 * np.arange(ps.nit)[:, None]       # (nit, 16)
                        sums = np.einsum("jsi,js->i", tile[:, g * 16:(g + 1) * 16, :].astype(np.float64), x[ks].astype(np.float64))
                        dst = (crit, defer, resid)[ps.dst]
                        dst[c * pl.CS + ps.owner[g], ps.dst_row[g]:ps.dst_row[g] + 4, r] = sums.astype(np.float32)
        return crit, defer, resid

    def run_teacher_forced(self, b):
        gc, pl, L, kw = self.gc, self.pl, self.L, self.kw
        NC, CS, P = pl.NC, pl.CS, pl.P
        G2, R, S, O = self.G2, self.R, self.S, self.O
        my, mx, ms, mo, qA, qB, qD, qS = pl.my, pl.mx, pl.ms, pl.mo, pl.qA, pl.qB, pl.qD, pl.qS
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

        def slices(vec_parts):
            """vec_parts: list of (published values per block [P][m], m); -> per-rank slices [c][m] concatenated"""
            res = []
            for r in range(CS):
                segs = []
                for vals, m in vec_parts:
                    segs.append(np.concatenate([vals[c * CS + r][:m] for c in range(NC)]))
                res.append(np.concatenate(segs))
            return res

        def x0_block(r):
            v = np.zeros(NC * mx, np.float32)
            for c in range(NC):
                base, cnt = own(R, NC, CS, c, r)
                v[c * mx:c * mx + cnt] = x0[base:base + cnt]
            return v

        for t in range(T):
            x0 = first_w @ x_tf[:, t] + first_b
            ypub = np.zeros((P, my), np.float32)
            xpub = np.zeros((P, mx), np.float32)
            for p in range(P):
                c, r = divmod(p, CS)
                base, cnt = own(R, NC, CS, c, r)
                xpub[p, :cnt] = x0[base:base + cnt]
            skipacc = np.zeros((P, ms), np.float32)
            for s_ in range(0, L):
                kind = K_FIRST if s_ == 0 else K_LAYER
                if s_ == 0:
                    xin = [np.concatenate([np.zeros(NC * my, np.float32), x0_block(r)]) for r in range(CS)]
                else:
                    xin = slices([(ypub, my), (xpub, mx)])
                    for r in range(CS):
                        assert len(xin[r]) == pl.Ky + pl.Kx
                crit, defer, resid = self.run_stage(kind, s_, xin)
                ynew = np.zeros((P, my), np.float32)
                xnew = np.zeros((P, mx), np.float32)
                for p in range(P):
                    c, r = divmod(p, CS)
                    bias = self.img[p]["b"]
                    y0, ny = own(G2, NC, CS, c, r)
                    pre = bias[pl.bo_zb + s_ * 4 * qA: pl.bo_zb + s_ * 4 * qA + 2 * my].copy()
                    for j in range(ny):
                        if gb is not None:
                            pre[2 * j] += gb[s_][y0 + j]
                            pre[2 * j + 1] += gb[s_][G2 + y0 + j]
                    if self.C:
                        cw = self.img[p]["cw"][s_ * qA * self.C * 4:(s_ + 1) * qA * self.C * 4].reshape(qA, self.C, 4)
                        cw = cw.transpose(0, 2, 1).reshape(4 * qA, self.C)[:2 * my]
                        pre += cw @ c_up[b, :, t].numpy()
                    for tap in range(kw - 1):
                        pre += rings[p][s_][tap][t % ((kw - 1 - tap) * dil[s_])]
                    z = crit[p, :2 * my].sum(axis=1) + pre
                    for j in range(ny):
                        ynew[p, j] = np.tanh(z[2 * j]) / (1.0 + np.exp(-z[2 * j + 1]))
                    if s_ >= 1:
                        x0r, nx = own(R, NC, CS, c, r)
                        o = resid[p, :mx].sum(axis=1) + bias[pl.bo_xb + s_ * 4 * qB: pl.bo_xb + s_ * 4 * qB + mx]
                        xnew[p, :nx] = ((o + xpub[p]) * rs2)[:nx]
                        layer = s_ - 1
                        for tap in range(kw - 1):
                            D = (kw - 1 - tap) * dil[layer]
                            rings[p][layer][tap][t % D] = defer[p, tap * 2 * my:(tap + 1) * 2 * my].sum(axis=1)
                        h = defer[p, 4 * qD:4 * qD + ms].sum(axis=1) + bias[pl.bo_sb + layer * 4 * qS: pl.bo_sb + layer * 4 * qS + ms]
                        skipacc[p] = h if layer == 0 else skipacc[p] + h
                ypub = ynew
                if s_ >= 1:
                    xpub = xnew
            # stage L: skip of the last layer
            xin = slices([(ypub, my), (xpub, mx)])
            crit, defer, _ = self.run_stage(K_TAIL, L, xin)
            skpub = np.zeros((P, ms), np.float32)
            for p in range(P):
                c, r = divmod(p, CS)
                bias = self.img[p]["b"]
                h = crit[p, :ms].sum(axis=1) + bias[pl.bo_sb + (L - 1) * 4 * qS: pl.bo_sb + (L - 1) * 4 * qS + ms]
                tot = h if L == 1 else skipacc[p] + h
                s0, ns = own(S, NC, CS, c, r)
                skpub[p, :ns] = np.maximum(tot * np.float32(math.sqrt(1.0 / L)), 0)[:ns]
                for tap in range(kw - 1):
                    D = (kw - 1 - tap) * dil[L - 1]
                    rings[p][L - 1][tap][t % D] = defer[p, tap * 2 * my:(tap + 1) * 2 * my].sum(axis=1)
            crit, _, _ = self.run_stage(K_HEAD1, L + 1, slices([(skpub, ms)]))
            h1pub = np.zeros((P, ms), np.float32)
            for p in range(P):
                c, r = divmod(p, CS)
                a0, na = own(S, NC, CS, c, r)
                h1pub[p, :na] = np.maximum(crit[p, :ms].sum(axis=1) + self.img[p]["b"][pl.bo_ha:pl.bo_ha + ms], 0)[:na]
            crit, _, _ = self.run_stage(K_HEAD2, L + 2, slices([(h1pub, ms)]))
            for p in range(P):
                c, r = divmod(p, CS)
                b0, nb = own(O, NC, CS, c, r)
                out[b0:b0 + nb, t] = (crit[p, :mo].sum(axis=1) + self.img[p]["b"][pl.bo_hb:pl.bo_hb + mo])[:nb]
        return out


@pytest.mark.parametrize("name,P,cluster", [("mol_cond", 5, 0), ("mol_cond", 32, 8), ("mol_cond", 16, 4),
                                            ("mulaw_softmax", 16, 0), ("gauss_speaker", 3, 0), ("mixgauss", 6, 2),
                                            ("mol_upsample", 24, 0), ("mol_cond", 16, 16)])
def test_packed_image_replays_reference(name, P, cluster):
    gc = GoldenCase(name)
    pm = PackedModel(gc, P, cluster)
    if cluster:
        assert pm.pl.CS == cluster
    got = pm.run_teacher_forced(0)
    ref = gc.arr["params_tf"][0]
    assert got.shape == ref.shape
    assert float(np.abs(got - ref).max()) <= 2e-5


def test_pass_lists_cover_every_row_once():
    """Every (owner, row slot) of every job receives exactly one tile per stage kind, critical passes precede
    deferred ones in every warp, and tiles do not overlap inside a blob."""
    cfg = make_config(layers=24, stacks=4, residual_channels=512, gate_channels=512, skip_out_channels=256,
                      out_channels=30, kernel_size=3, cin_channels=80, gin_channels=-1, scalar_input=True,
                      output_distribution="Logistic")
    for batch in (1, 8):
        pl, passes = N.plan_passes(cfg, batch)
        assert (pl.NC, pl.CS, pl.P, pl.BT) == (16, 8, 128, batch)
        for kind in range(5):
            seen = {}
            spans = []
            for wv in range(8):
                b0, n, nc = pl.pass_begin[kind][wv], pl.pass_count[kind][wv], pl.pass_crit[kind][wv]
                for i, ps in enumerate(passes[b0:b0 + n]):
                    assert (ps.dst != 1) == (i < nc)
                    spans.append((ps.w_off, ps.w_off + ps.nit * 128))
                    for g in range(2):
                        if ps.owner[g] >= 0:
                            key = (ps.dst, ps.owner[g], ps.dst_row[g])
                            assert key not in seen
                            seen[key] = ps.job
            rows_c = {(o, r) for (d, o, r) in seen if d == 0}
            rows_d = {(o, r) for (d, o, r) in seen if d == 1}
            rows_x = {(o, r) for (d, o, r) in seen if d == 2}
            assert len(rows_c) * 4 == pl.rows_c[kind] * pl.CS and len(rows_d) * 4 == pl.rows_d[kind] * pl.CS
            assert len(rows_x) * 4 == (pl.rows_x * pl.CS if kind == K_LAYER else 0)
            if kind in (K_FIRST, K_LAYER):
                spans.sort()
                assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
