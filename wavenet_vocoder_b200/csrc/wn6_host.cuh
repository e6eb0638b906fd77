// wn6_host.cuh — planner and weight packer of the cluster kernel (included by wn_host.cu).
// Pure host arithmetic: exercised without a GPU through wn_plan_only / wn_pack_cta / wn_plan_passes.
#pragma once
#include "wn6_plan.h"

static int align_up6(long long v, int a) { return (int)(((v + a - 1) / a) * a); }

// default number of co-resident clusters of size cs on a GPU with `num_sms` SMs (B200: 8 GPCs of 16-20 SMs,
// a cluster lives inside one GPC); wn_create replaces it by cudaOccupancyMaxActiveClusters
static int default_max_clusters(int num_sms, int cs) {
    if (cs >= 16) return std::max(1, std::min(num_sms / 16, 8));
    if (cs == 8) return std::max(1, std::min(num_sms / 8, 16));
    if (cs == 4) return std::max(1, (num_sms * 8 / 9) / 4);
    return std::max(1, num_sms / cs);
}

struct JobSpec {
    int job, quads_per_owner, x_off, klen, dst, row_base;
};

static int32_t build_plan6(const wn_config& c, int batch, int num_sms, long long smem_cap, int max_clusters_hint,
                           Wn6Plan& pl, std::vector<Wn6Pass>& passes, std::vector<int>& ringtab) {
    memset(&pl, 0, sizeof(pl));
    passes.clear();
    if (c.abi_version != WN_ABI_VERSION) return fail(WN_ERR_INVALID, "wn_config.abi_version mismatch");
    if (c.layers < 1 || c.stacks < 1 || c.layers % c.stacks != 0)
        return fail(WN_ERR_INVALID, "layers must be a positive multiple of stacks (wavenet.py:117)");
    if (c.gate_channels < 2 || (c.gate_channels & 1)) return fail(WN_ERR_INVALID, "gate_channels must be even");
    if (c.kernel_size < 1 || c.kernel_size > 8) return fail(WN_ERR_INVALID, "kernel_size out of range [1,8]");
    if (c.residual_channels < 1 || c.skip_channels < 1 || c.out_channels < 1)
        return fail(WN_ERR_INVALID, "channel counts must be positive");
    if (c.residual_channels > 8192 || c.gate_channels > 16384 || c.skip_channels > 8192 || c.out_channels > 8192)
        return fail(WN_ERR_INVALID, "channel count too large (unsupported shape)");
    if (c.cin_channels < 0 || c.cin_channels > 32 * WN6_MAX_CI)
        return fail(WN_ERR_INVALID, "cin_channels must be in [0,128]");
    if (c.gin_channels < 0) return fail(WN_ERR_INVALID, "gin_channels must be >= 0");
    if (c.layers / c.stacks > 20) return fail(WN_ERR_INVALID, "dilation 2^(layers/stacks) too large");
    if (c.input_kind == WN_INPUT_SCALAR) {
        if (c.head_kind == WN_HEAD_MOL) {
            if (c.out_channels % 3 != 0) return fail(WN_ERR_INVALID, "MoL head needs out_channels % 3 == 0 (mixture.py:130)");
        } else if (c.head_kind == WN_HEAD_GAUSS) {
            if (c.out_channels != 2 && c.out_channels % 3 != 0)
                return fail(WN_ERR_INVALID, "Gaussian head needs out_channels == 2 or % 3 == 0 (mixture.py:229-234)");
        } else
            return fail(WN_ERR_INVALID, "scalar input needs a MoL or Gaussian head (wavenet.py:322-330)");
    } else if (c.input_kind == WN_INPUT_ONEHOT) {
        if (c.head_kind != WN_HEAD_SOFTMAX) return fail(WN_ERR_INVALID, "one-hot input needs the softmax head");
    } else
        return fail(WN_ERR_INVALID, "bad input_kind");
    if (batch < 1) return fail(WN_ERR_INVALID, "batch must be >= 1");
    if (num_sms < 1) return fail(WN_ERR_INVALID, "no SMs");

    pl.L = c.layers;
    pl.per_stack = c.layers / c.stacks;
    pl.R = c.residual_channels;
    pl.G = c.gate_channels;
    pl.G2 = c.gate_channels / 2;
    pl.S = c.skip_channels;
    pl.O = c.out_channels;
    pl.kw = c.kernel_size;
    pl.C = c.cin_channels;
    pl.gin = c.gin_channels;
    pl.input_kind = c.input_kind;
    pl.head_kind = c.head_kind;
    pl.Kmix = (c.head_kind == WN_HEAD_SOFTMAX) ? 0 : (c.out_channels == 2 ? 1 : c.out_channels / 3);
    pl.skip_scale = (float)sqrt(1.0 / (double)c.layers);
    pl.BT = batch <= 1 ? 1 : (batch <= 2 ? 2 : (batch <= 4 ? 4 : 8));
    pl.NS = pl.L + 3;
    const int BT = pl.BT;

    // ---- grid: NC clusters of CS blocks
    int CS = c.cluster_size > 0 ? c.cluster_size : env_int("WN_CLUSTER", 8);
    if (CS < 1 || CS > WN6_MAX_CS || (CS & (CS - 1))) return fail(WN_ERR_INVALID, "cluster size must be a power of two in [1,16]");
    int P = c.num_ctas > 0 ? c.num_ctas : env_int("WN_NUM_CTAS", 0);
    int NC;
    if (P > 0) {
        if (P > num_sms) return fail(WN_ERR_INVALID, "num_ctas exceeds the SM count (blocks must be co-resident)");
        while (CS > 1 && (P % CS) != 0) CS >>= 1;
        NC = P / CS;
    } else {
        const int maxc = max_clusters_hint > 0 ? max_clusters_hint : default_max_clusters(num_sms, CS);
        // shrink the cluster for tiny models so that the ranks still own rows
        while (CS > 1 && pl.G2 < CS) CS >>= 1;
        NC = std::max(1, std::min(maxc, pl.G2 / CS));
        if (NC * CS > num_sms) NC = std::max(1, num_sms / CS);
    }
    pl.NC = NC;
    pl.CS = CS;
    pl.P = NC * CS;
    auto per_block = [&](int rows) { return wn6_ceil_div(wn6_ceil_div(rows, NC), CS); };
    pl.my = per_block(pl.G2);
    pl.mx = per_block(pl.R);
    pl.ms = per_block(pl.S);
    pl.mo = per_block(pl.O);
    pl.qA = wn6_ceil_div(2 * pl.my, 4);
    pl.qB = wn6_ceil_div(pl.mx, 4);
    pl.qD = wn6_ceil_div((pl.kw - 1) * 2 * pl.my, 4);
    pl.qS = wn6_ceil_div(pl.ms, 4);
    pl.qHA = pl.qS;
    pl.qHB = wn6_ceil_div(pl.mo, 4);
    pl.Ky = NC * pl.my;
    pl.Kx = NC * pl.mx;
    pl.Ksk = NC * pl.ms;
    pl.Kh2 = NC * pl.mo;
    if (pl.Ky + pl.Kx > 30000) return fail(WN_ERR_INVALID, "K-slice too long");

    // ---- partial buffers
    pl.rows_c[WN6_K_FIRST] = 4 * pl.qA;
    pl.rows_c[WN6_K_LAYER] = 4 * pl.qA;
    pl.rows_x = 4 * pl.qB;
    pl.nrow_x = 4 * pl.qB;
    pl.rows_c[WN6_K_TAIL] = 4 * pl.qS;
    pl.rows_c[WN6_K_HEAD1] = 4 * pl.qHA;
    pl.rows_c[WN6_K_HEAD2] = 4 * pl.qHB;
    pl.rows_d[WN6_K_LAYER] = 4 * pl.qD + 4 * pl.qS;
    pl.rows_d[WN6_K_TAIL] = 4 * pl.qD;
    pl.nrow_c = 0;
    for (int k = 0; k < WN6_NKIND; ++k) pl.nrow_c = std::max(pl.nrow_c, pl.rows_c[k]);
    pl.nrow_d = std::max(4, 4 * pl.qD + 4 * pl.qS);
    if ((long long)pl.nrow_c * CS * BT * 4 >= (1 << 20) || (long long)pl.nrow_d * CS * BT * 4 >= (1 << 20))
        return fail(WN_ERR_INVALID, "too many rows per block for the mbarrier transaction count (use more blocks)");

    // ---- passes.  Jobs of a stage kind in order (critical first); every job's quads are dealt two per pass,
    // passes round-robin over the compute warps continuing across jobs.
    const int Ky = pl.Ky, Kx = pl.Kx;
    std::vector<std::vector<JobSpec>> kinds(WN6_NKIND);
    kinds[WN6_K_FIRST] = {{WN6_J_A0, pl.qA, Ky, Kx, 0, 0}};
    kinds[WN6_K_LAYER] = {{WN6_J_A, pl.qA, 0, Ky + Kx, 0, 0},
                          {WN6_J_B, pl.qB, 0, Ky, 2, 0},
                          {WN6_J_D, pl.qD, Ky, Kx, 1, 0},
                          {WN6_J_S, pl.qS, 0, Ky, 1, 4 * pl.qD}};
    kinds[WN6_K_TAIL] = {{WN6_J_S, pl.qS, 0, Ky, 0, 0}, {WN6_J_D, pl.qD, Ky, Kx, 1, 0}};
    kinds[WN6_K_HEAD1] = {{WN6_J_HA, pl.qHA, 0, pl.Ksk, 0, 0}};
    kinds[WN6_K_HEAD2] = {{WN6_J_HB, pl.qHB, 0, pl.Ksk, 0, 0}};
    int xin_vals = 16;
    // blob floats accumulate per blob: first = FIRST; layer = LAYER; tail = TAIL + HEAD1 + HEAD2
    int blob_fill[3] = {0, 0, 0};
    auto blob_of_kind = [](int k) { return k == WN6_K_FIRST ? 0 : (k == WN6_K_LAYER ? 1 : 2); };
    for (int k = 0; k < WN6_NKIND; ++k) {
        std::vector<std::vector<Wn6Pass>> per_warp(WN6_NCW);
        std::vector<int> ncrit(WN6_NCW, 0);
        int rr = 0;
        for (const JobSpec& js : kinds[k]) {
            const int nquads = CS * js.quads_per_owner;
            if (nquads == 0 || js.klen == 0) continue;
            const int nit = wn6_ceil_div(js.klen, WN6_TPQ);
            xin_vals = std::max(xin_vals, js.x_off + WN6_TPQ * nit);
            for (int q = 0; q < nquads; q += 2) {
                Wn6Pass ps;
                memset(&ps, 0, sizeof(ps));
                ps.nit = (int16_t)nit;
                ps.x_off = (int16_t)js.x_off;
                ps.dst = (int8_t)js.dst;
                ps.job = (int8_t)js.job;
                for (int g = 0; g < 2; ++g) {
                    const int qq = q + g;
                    if (qq < nquads) {
                        ps.owner[g] = (int8_t)(qq / js.quads_per_owner);
                        ps.dst_row[g] = (int16_t)(js.row_base + (qq % js.quads_per_owner) * 4);
                        ps.quad[g] = (int16_t)qq;
                    } else {
                        ps.owner[g] = -1;
                        ps.dst_row[g] = 0;
                        ps.quad[g] = -1;
                    }
                }
                const int w = rr % WN6_NCW;
                ++rr;
                per_warp[w].push_back(ps);
                if (js.dst != 1) ncrit[w] = (int)per_warp[w].size();
            }
        }
        // a warp's critical passes must precede its deferred ones (jobs are listed in that order already)
        for (int w = 0; w < WN6_NCW; ++w) {
            pl.pass_begin[k][w] = (int)passes.size();
            pl.pass_count[k][w] = (int)per_warp[w].size();
            pl.pass_crit[k][w] = ncrit[w];
            for (Wn6Pass& ps : per_warp[w]) {
                int& fill = blob_fill[blob_of_kind(k)];
                ps.w_off = fill;
                fill += ps.nit * 32 * 4;
                passes.push_back(ps);
            }
        }
    }
    pl.npass = (int)passes.size();
    if (pl.npass > 4096) return fail(WN_ERR_INVALID, "too many passes (unsupported shape)");
    pl.xin_vals = align_up6(xin_vals, 4);
    pl.fb_floats = align_up6(blob_fill[0], 4);
    pl.lb_floats = align_up6(blob_fill[1], 4);
    pl.tb_floats = align_up6(blob_fill[2], 4);
    pl.slot_floats = align_up6(std::max(pl.fb_floats, std::max(pl.L > 1 ? pl.lb_floats : 0, pl.tb_floats)), 32);
    pl.cta_w_floats = (long long)pl.fb_floats + (long long)(pl.L - 1) * pl.lb_floats + pl.tb_floats;
    pl.nblobs = pl.L + 1;
    pl.cta_cw_floats = (long long)pl.L * pl.qA * pl.C * 4;
    // ---- biases
    int bo = 0;
    pl.bo_zb = bo; bo += pl.L * 4 * pl.qA;
    pl.bo_xb = bo; bo += pl.L * 4 * pl.qB;
    pl.bo_sb = bo; bo += pl.L * 4 * pl.qS;
    pl.bo_ha = bo; bo += 4 * pl.qHA;
    pl.bo_hb = bo; bo += 4 * pl.qHB;
    pl.cta_b_floats = align_up6(bo, 4);

    // ---- exchange map (pairs); slices start on 256-byte boundaries
    pl.rs_yx = align_up6((long long)(Ky + Kx) * BT, 32);
    pl.rs_sk = align_up6((long long)pl.Ksk * BT, 32);
    pl.rs_h2 = align_up6((long long)pl.Kh2 * BT, 32);
    pl.ex_yx = 0;
    pl.ex_sk = (long long)pl.L * CS * pl.rs_yx;
    pl.ex_h1 = pl.ex_sk + (long long)CS * pl.rs_sk;
    pl.ex_h2 = pl.ex_h1 + (long long)CS * pl.rs_sk;
    pl.ex_pairs = pl.ex_h2 + (long long)CS * pl.rs_h2 + 32;

    // ---- history rings: tap k (0 = oldest) is consumed (kw-1-k)*d steps later
    ringtab.assign((size_t)pl.L * std::max(pl.kw - 1, 0) * 2, 0);
    long long pos = 0;
    for (int l = 0; l < pl.L; ++l)
        for (int k = 0; k < pl.kw - 1; ++k) {
            const int D = (pl.kw - 1 - k) * wn6_dilation(pl, l);
            ringtab[((size_t)l * (pl.kw - 1) + k) * 2] = (int)pos;
            ringtab[((size_t)l * (pl.kw - 1) + k) * 2 + 1] = D;
            pos += D;
        }
    pl.ring_pos_total = pos;
    const long long ring_bytes = pos * 4 * pl.qA * BT * 4;

    // ---- shared memory map
    auto layout = [&](bool ring_smem) -> long long {
        long long off = 0;
        auto take = [&](long long bytes, int al) {
            off = ((off + al - 1) / al) * al;
            long long r = off;
            off += bytes;
            return (int)r;
        };
        const long long tab = (long long)pl.L * 4 * pl.qA * BT * 4;
        pl.sm_bar = take((long long)(2 * pl.nblobs + 24) * 8, 16);
        pl.sm_misc = take(16, 16);
        pl.sm_in = take((long long)BT * 8 + (pl.input_kind == WN_INPUT_ONEHOT ? (long long)BT * pl.O * 4 : 0), 16);
        pl.sm_pass = take((long long)pl.npass * (long long)sizeof(Wn6Pass), 16);
        pl.sm_ringtab = take((long long)ringtab.size() / 2 * 3 * 4 + 16, 16);
        pl.sm_xin = take(2LL * pl.xin_vals * BT * 4, 16);
        pl.sm_part = take(2LL * pl.nrow_c * CS * BT * 4, 16);
        pl.sm_dpart = take(2LL * pl.nrow_d * CS * BT * 4, 16);
        pl.sm_partx = take(2LL * pl.nrow_x * CS * BT * 4, 16);
        pl.sm_sb = take(tab, 16);
        pl.sm_pre = take(tab, 16);
        pl.sm_cond = take(pl.C > 0 ? 2 * tab : 16, 16);
        pl.sm_bias = take((long long)pl.cta_b_floats * 4, 16);
        pl.sm_skipacc = take((long long)4 * pl.qS * BT * 4, 16);
        pl.sm_xown = take((long long)8 * pl.qB * BT * 4, 16);
        pl.sm_hs = take((long long)pl.O * BT * 4 + (long long)CS * pl.Kh2 * BT * 4, 16);
        pl.sm_noise = take((long long)BT * (pl.O + 2) * 4, 16);
        pl.sm_x0w = take((long long)(2 * Kx + 8 * pl.qB) * 4, 16);
        pl.sm_ring = take(ring_smem ? ring_bytes : 16, 16);
        pl.sm_slots = take(0, 128);
        return off;
    };
    const long long slot_bytes = (long long)pl.slot_floats * 4;
    const int want_ring_smem = env_int("WN_RING_SMEM", -1);
    bool ring_smem = (want_ring_smem != 0) && ring_bytes <= 96 * 1024;
    long long fixed = layout(ring_smem);
    long long fit = (smem_cap - fixed) / slot_bytes;
    if (ring_smem && want_ring_smem < 0 && fit < std::min<long long>(pl.nblobs, 3)) {
        ring_smem = false;
        fixed = layout(false);
        fit = (smem_cap - fixed) / slot_bytes;
    }
    pl.ring_in_smem = ring_smem ? 1 : 0;
    if (fit >= pl.nblobs) {
        pl.nres = pl.nblobs;
        pl.nring = 0;
    } else {
        if (fit < 2) return fail(WN_ERR_INVALID, "shared memory too small for two weight slots (use more blocks)");
        int nr = c.ring_slots > 0 ? c.ring_slots : env_int("WN_RING_SLOTS", 4);
        nr = (int)std::max<long long>(2, std::min<long long>(nr, fit));
        pl.nring = nr;
        pl.nres = (int)fit - nr;
        const int force_res = env_int("WN_RESIDENT", -1);
        if (force_res >= 0 && force_res < pl.nres) pl.nres = force_res;
    }
    pl.smem_bytes = (int)(pl.sm_slots + (long long)(pl.nres + pl.nring) * slot_bytes);
    if (pl.smem_bytes > smem_cap) return fail(WN_ERR_INVALID, "shared memory map exceeds the per-block limit");
    return WN_OK;
}

// ------------------------------------------------------------------------------------------
// packer
// ------------------------------------------------------------------------------------------
// matrix element of `job` for the stage that uses layer `layer` (the layer whose gates stage `layer` evaluates):
// row = (owner o, row-in-owner ro) of cluster c, column = entry k of rank r's K-slice.  Returns 0 for padding.
struct Pack6 {
    const Wn6Plan& pl;
    const wn_weights& w;
    const Folded& f;
    int c, r;
    // global gate row of (owner, ro): pair i = ro/2, half ab = ro&1
    bool gate_row(int o, int ro, int& grow) const {
        int base, cnt;
        wn6_own(pl.G2, pl.NC, pl.CS, c, o, base, cnt);
        const int i = ro >> 1;
        if (i >= cnt) return false;
        grow = (ro & 1) ? pl.G2 + base + i : base + i;
        return true;
    }
    bool plain_row(int rows, int o, int ro, int& row) const {
        int base, cnt;
        wn6_own(rows, pl.NC, pl.CS, c, o, base, cnt);
        if (ro >= cnt) return false;
        row = base + ro;
        return true;
    }
    int ycol(int k) const { return (k >= 0 && k < pl.Ky) ? wn6_slice_index(pl.G2, pl.NC, pl.CS, pl.my, r, k) : -1; }
    int xcol(int k) const { return (k >= pl.Ky && k < pl.Ky + pl.Kx) ? wn6_slice_index(pl.R, pl.NC, pl.CS, pl.mx, r, k - pl.Ky) : -1; }
    int scol(int k) const { return (k >= 0 && k < pl.Ksk) ? wn6_slice_index(pl.S, pl.NC, pl.CS, pl.ms, r, k) : -1; }

    // `stage` = the stage the pass belongs to (0..L+2)
    float elem(int job, int stage, int o, int ro, int k) const {
        const int R = pl.R, G2 = pl.G2, kw = pl.kw, S = pl.S;
        int row, col;
        switch (job) {
            case WN6_J_A0:
                if (!gate_row(o, ro, row) || (col = xcol(k)) < 0) return 0.f;
                return f.V[0][(size_t)row * R + col];
            case WN6_J_A:
                if (!gate_row(o, ro, row)) return 0.f;
                if ((col = ycol(k)) >= 0) return f.M[stage - 1][(size_t)row * G2 + col];
                if ((col = xcol(k)) >= 0) return f.V[stage][(size_t)row * R + col];
                return 0.f;
            case WN6_J_B:
                if (!plain_row(R, o, ro, row) || (col = ycol(k)) < 0) return 0.f;
                return w.layers[stage - 1].out_w[(size_t)row * G2 + col];
            case WN6_J_D: {
                const int tap = ro / (2 * pl.my), rr = ro % (2 * pl.my);
                if (tap >= kw - 1 || !gate_row(o, rr, row) || (col = xcol(k)) < 0) return 0.f;
                return w.layers[stage - 1].conv_w[(size_t)row * kw * R + (size_t)tap * R + col];   // conv.py:56-61: col = k*R + r
            }
            case WN6_J_S:
                if (!plain_row(S, o, ro, row) || (col = ycol(k)) < 0) return 0.f;
                return w.layers[stage - 1].skip_w[(size_t)row * G2 + col];
            case WN6_J_HA:
                if (!plain_row(S, o, ro, row) || (col = scol(k)) < 0) return 0.f;
                return w.last_a_w[(size_t)row * S + col];
            case WN6_J_HB:
                if (!plain_row(pl.O, o, ro, row) || (col = scol(k)) < 0) return 0.f;
                return w.last_b_w[(size_t)row * S + col];
        }
        return 0.f;
    }
};

static int quads_per_owner6(const Wn6Plan& pl, int job) {
    switch (job) {
        case WN6_J_A0: case WN6_J_A: return pl.qA;
        case WN6_J_B: return pl.qB;
        case WN6_J_D: return pl.qD;
        case WN6_J_S: return pl.qS;
        case WN6_J_HA: return pl.qHA;
        default: return pl.qHB;
    }
}

// packed image of block `p`: first blob, L-1 layer blobs, tail blob; tiles in pass order, [j][lane][4 rows]
static void pack6_cta(const Wn6Plan& pl, const std::vector<Wn6Pass>& passes, const wn_weights& w, const Folded& f, int p,
                      float* out) {
    memset(out, 0, (size_t)pl.cta_w_floats * sizeof(float));
    Pack6 pk{pl, w, f, p / pl.CS, p % pl.CS};
    auto pack_kind = [&](int kind, int stage, float* blob) {
        for (int wv = 0; wv < WN6_NCW; ++wv)
            for (int i = 0; i < pl.pass_count[kind][wv]; ++i) {
                const Wn6Pass& ps = passes[pl.pass_begin[kind][wv] + i];
                const int qpo = quads_per_owner6(pl, ps.job);
                float* tile = blob + ps.w_off;
                for (int g = 0; g < 2; ++g) {
                    if (ps.owner[g] < 0) continue;
                    const int o = ps.quad[g] / qpo, ql = ps.quad[g] % qpo;
                    for (int j = 0; j < ps.nit; ++j)
                        for (int sub = 0; sub < WN6_TPQ; ++sub) {
                            const int k = ps.x_off + sub + WN6_TPQ * j;
                            float* dst = tile + ((size_t)j * 32 + g * 16 + sub) * 4;
                            for (int i4 = 0; i4 < 4; ++i4) dst[i4] = pk.elem(ps.job, stage, o, ql * 4 + i4, k);
                        }
                }
            }
    };
    pack_kind(WN6_K_FIRST, 0, out);
    for (int s = 1; s < pl.L; ++s) pack_kind(WN6_K_LAYER, s, out + wn6_blob_off(pl, s));
    float* tb = out + wn6_blob_off(pl, pl.L);
    pack_kind(WN6_K_TAIL, pl.L, tb);
    pack_kind(WN6_K_HEAD1, pl.L + 1, tb);
    pack_kind(WN6_K_HEAD2, pl.L + 2, tb);
}

// biases of the rows block p owns
static void pack6_bias(const Wn6Plan& pl, const wn_weights& w, const Folded& f, int p, float* out) {
    memset(out, 0, (size_t)pl.cta_b_floats * sizeof(float));
    Pack6 pk{pl, w, f, p / pl.CS, p % pl.CS};
    const int o = pk.r;
    int row;
    for (int l = 0; l < pl.L; ++l) {
        for (int ro = 0; ro < 2 * pl.my; ++ro)
            if (pk.gate_row(o, ro, row)) out[pl.bo_zb + l * 4 * pl.qA + ro] = f.zb[l][row];
        if (l >= 1)
            for (int ro = 0; ro < pl.mx; ++ro)
                if (pk.plain_row(pl.R, o, ro, row))
                    out[pl.bo_xb + l * 4 * pl.qB + ro] = w.layers[l - 1].out_b ? w.layers[l - 1].out_b[row] : 0.f;
        for (int ro = 0; ro < pl.ms; ++ro)
            if (pk.plain_row(pl.S, o, ro, row))
                out[pl.bo_sb + l * 4 * pl.qS + ro] = w.layers[l].skip_b ? w.layers[l].skip_b[row] : 0.f;
    }
    for (int ro = 0; ro < pl.ms; ++ro)
        if (pk.plain_row(pl.S, o, ro, row)) out[pl.bo_ha + ro] = w.last_a_b ? w.last_a_b[row] : 0.f;
    for (int ro = 0; ro < pl.mo; ++ro)
        if (pk.plain_row(pl.O, o, ro, row)) out[pl.bo_hb + ro] = w.last_b_b ? w.last_b_b[row] : 0.f;
}

// conditioning rows of block p: [L][qA][C][4]
static void pack6_cw(const Wn6Plan& pl, const wn_weights& w, int p, float* out) {
    if (pl.C <= 0) return;
    memset(out, 0, (size_t)pl.cta_cw_floats * sizeof(float));
    Folded dummy;
    Pack6 pk{pl, w, dummy, p / pl.CS, p % pl.CS};
    int row;
    for (int l = 0; l < pl.L; ++l) {
        const float* cwm = w.layers[l].cond_w;
        float* grp = out + (size_t)l * pl.qA * pl.C * 4;
        for (int ro = 0; ro < 2 * pl.my; ++ro)
            if (pk.gate_row(pk.r, ro, row))
                for (int ch = 0; ch < pl.C; ++ch) grp[((size_t)(ro >> 2) * pl.C + ch) * 4 + (ro & 3)] = cwm[(size_t)row * pl.C + ch];
    }
}
