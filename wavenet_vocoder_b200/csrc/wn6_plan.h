// wn6_plan.h — execution plan of the cluster kernel (wn6_kernel.cuh), shared by the host
// planner/packer and the device code.
//
// One generated sample is a chain of L+3 small matrix-vector "stages" with a strict serial dependency
// (reference wavenet.py:296-336).  The grid is NC thread-block clusters of CS blocks (P = NC*CS, one
// block per SM).  Every stage is decomposed in TWO dimensions:
//   * cluster c owns a block of OUTPUT rows of every matrix; inside the cluster the rows are dealt to
//     the CS "owner" blocks that finalise them (bias, gate, residual) and publish the results;
//   * rank r of every cluster owns the K-slice of the stage INPUT made of the values finalised by the
//     rank-r blocks of all clusters.
// Per stage a block therefore (1) polls only its K/CS slice of the tagged (value, tag) pairs in L2 --
// every pair is polled by NC blocks instead of P --, (2) multiplies it with its [cluster rows x K-slice]
// weight tile from shared memory, (3) sends the partial sums to the owner of every row through
// distributed shared memory (st.async + mbarrier complete_tx), and (4) as an owner sums CS partials,
// applies the gate / residual / ReLU and publishes its few values for the next stage.
//
// Work inside a block is described by "passes": one warp, two row quads (lanes 0-15 / 16-31), nit
// k-steps; lane `sub` of a quad handles k = x_off + sub + 16*j.  The planner builds the pass lists (critical
// jobs first, then the deferred ones: queued older-tap products and skip rows); the packer lays the weights
// out in exactly the order the lanes read them ([pass][j][lane][4 rows], one conflict-free 16-byte
// shared-memory load per lane and k-step).  All of it is plain arithmetic on the model shape, so the host
// tests replay it without a GPU.
#pragma once
#include <stdint.h>

#ifndef WN_HD
#ifdef __CUDACC__
#define WN_HD __host__ __device__ __forceinline__
#else
#define WN_HD inline
#endif
#endif

#define WN6_TPQ 16                // lanes per row quad
#define WN6_NPW 2                 // polling warps (also run the sampler)
#define WN6_NCW 8                 // compute warps
#define WN6_NFW 2                 // finaliser warps: F0 gate / skip / head items, F1 residual items
#define WN6_W_POLL 0
#define WN6_W_COMP (WN6_NPW)
#define WN6_W_F0 (WN6_NPW + WN6_NCW)
#define WN6_W_F1 (WN6_W_F0 + 1)
#define WN6_W_DF (WN6_W_F0 + 2)   // deferred finaliser: history rings, skip accumulator, pre-sum table
#define WN6_W_TMA (WN6_W_F0 + 3)  // weight streaming
#define WN6_W_COND (WN6_W_F0 + 4) // local-conditioning projection, one step ahead
#define WN6_NWARPS (WN6_W_F0 + 5)
#define WN6_NTHREADS (32 * WN6_NWARPS)
#define WN6_MAX_BT 8
#define WN6_MAX_CI 4              // local-conditioning channels <= 32*WN6_MAX_CI
#define WN6_MAX_CS 16

// stage kinds (which pass list / finalisation a stage uses)
enum { WN6_K_FIRST = 0, WN6_K_LAYER = 1, WN6_K_TAIL = 2, WN6_K_HEAD1 = 3, WN6_K_HEAD2 = 4, WN6_NKIND = 5 };
// jobs (which matrix a pass multiplies)
enum { WN6_J_A0 = 0,   // stage 0: current tap of layer 0 x x_0                      (critical)
       WN6_J_A = 1,    // [M_{s-1} | V_s] x (y_{s-1}, x_{s-1}): gate pre-activations (critical)
       WN6_J_B = 2,    // conv1x1_out_{s-1} x y_{s-1}: the residual stream x_s        (critical)
       WN6_J_D = 3,    // older taps of layer s-1 x x_{s-1}: queued for t+d, t+2d ..  (deferred)
       WN6_J_S = 4,    // conv1x1_skip_{s-1} x y_{s-1}   (deferred; critical in the tail stage)
       WN6_J_HA = 5,   // last_conv_layers[1] x relu(skip)
       WN6_J_HB = 6,   // last_conv_layers[3] x relu(h1)
       WN6_NJOB = 7 };

struct Wn6Pass {
    int32_t w_off;        // float offset of the tile inside the stage's blob: [nit][32 lanes][4 rows]
    int16_t nit;          // k-steps
    int16_t x_off;        // first k of the tile in the stage input slice
    int16_t dst_row[2];   // per lane group: first row slot in the owner's partial buffer
    int8_t owner[2];      // per lane group: owner rank inside the cluster, -1 = idle quad
    int8_t dst;           // partial buffer of the owner: 0 = F0 (gate / skip / head rows), 1 = DF (deferred rows),
                          // 2 = F1 (residual rows)
    int8_t job;           // WN6_J_*  (packer / tests only)
    int16_t quad[2];      // per lane group: quad index inside the job (packer / tests only)
};

struct Wn6Plan {
    // ---- model shape (wavenet.py:98-111)
    int L, per_stack, R, G, G2, S, O, kw, C, gin, input_kind, head_kind, Kmix;
    float skip_scale;           // sqrt(1/L), wavenet.py:313
    // ---- grid
    int NC, CS, P, BT;
    // ---- values finalised per block (uniform maxima; the last ranks / clusters may own fewer)
    int my, mx, ms, mo;         // gate pairs, residual rows, skip (= head-1) rows, head-2 rows
    // ---- row quads per owner and job
    int qA, qB, qD, qS, qHA, qHB;
    // ---- K-slices (values per rank): y part, x part, skip / head-1 vector, head-2 vector
    int Ky, Kx, Ksk, Kh2;
    int xin_vals;               // values (x BT floats) of one stage-input buffer
    // ---- exchange in L2 (offsets / strides in pairs)
    int NS;                     // stages per step = L+3
    int rs_yx, rs_sk, rs_h2;    // per-rank stride of a slice
    long long ex_yx, ex_sk, ex_h1, ex_h2, ex_pairs;
    // ---- partial buffers of an owner (rows x CS x BT floats, double buffered): one per finaliser warp, each
    // with its own mbarrier pair, so that every barrier is armed and waited on by exactly one warp
    int nrow_c, nrow_d, nrow_x;
    int rows_c[WN6_NKIND];      // rows sent to F0 per stage kind (tx bytes = rows*CS*BT*4)
    int rows_d[WN6_NKIND];      // rows sent to DF per stage kind
    int rows_x;                 // rows sent to F1 in a layer stage
    // ---- pass lists: pass_begin[kind][warp] .. +pass_count, the first pass_crit of them critical
    int npass;
    int pass_begin[WN6_NKIND][WN6_NCW], pass_count[WN6_NKIND][WN6_NCW], pass_crit[WN6_NKIND][WN6_NCW];
    // ---- blobs (floats): first (stage 0), layer (stages 1..L-1), tail (stages L, L+1, L+2)
    int fb_floats, lb_floats, tb_floats, slot_floats;
    long long cta_w_floats;     // packed floats per block = fb + (L-1)*lb + tb
    int nblobs, nres, nring;
    // ---- biases of an owner's rows: [L][4qA] zb | [L][4qB] xb | [L][4qS] sb | [4qHA] | [4qHB]
    int bo_zb, bo_xb, bo_sb, bo_ha, bo_hb, cta_b_floats;
    // ---- conditioning rows of an owner, read from L2 by the conditioning warp: [L][qA][C][4]
    long long cta_cw_floats;
    // ---- history rings of the older-tap products: one position = 4qA*BT floats
    int ring_in_smem;
    long long ring_pos_total;
    // ---- shared memory map (byte offsets)
    int sm_bar, sm_misc, sm_pass, sm_ringtab, sm_xin, sm_part, sm_dpart, sm_partx, sm_sb, sm_pre, sm_cond, sm_bias,
        sm_skipacc, sm_xown, sm_hs, sm_noise, sm_in, sm_x0w, sm_ring, sm_slots, smem_bytes;
};

// balanced split of `rows` over n parts: part p owns [base, base+cnt)
WN_HD void wn6_part(int rows, int n, int p, int& base, int& cnt) {
    const int q = rows / n, r = rows % n;
    base = p * q + (p < r ? p : r);
    cnt = q + (p < r ? 1 : 0);
}
// rows of a matrix with `rows` rows that block (c, r) finalises: global [base, base+cnt)
WN_HD void wn6_own(int rows, int NC, int CS, int c, int r, int& base, int& cnt) {
    int cb, cc;
    wn6_part(rows, NC, c, cb, cc);
    int ob, oc;
    wn6_part(cc, CS, r, ob, oc);
    base = cb + ob;
    cnt = oc;
}
// entry k of rank r's K-slice of a vector with `rows` entries and `m` slots per block: global index or -1
WN_HD int wn6_slice_index(int rows, int NC, int CS, int m, int r, int k) {
    const int c = k / m, i = k % m;
    if (c >= NC) return -1;
    int base, cnt;
    wn6_own(rows, NC, CS, c, r, base, cnt);
    return i < cnt ? base + i : -1;
}
WN_HD int wn6_ceil_div(int a, int b) { return (a + b - 1) / b; }
WN_HD int wn6_dilation(const Wn6Plan& pl, int l) { return 1 << (l % pl.per_stack); }
// stage -> kind, blob, exchange it publishes
WN_HD int wn6_kind(const Wn6Plan& pl, int s) {
    return s == 0 ? WN6_K_FIRST : (s < pl.L ? WN6_K_LAYER : (s == pl.L ? WN6_K_TAIL : (s == pl.L + 1 ? WN6_K_HEAD1 : WN6_K_HEAD2)));
}
WN_HD int wn6_blob_of_stage(const Wn6Plan& pl, int s) { return s < pl.L ? s : pl.L; }
WN_HD long long wn6_blob_off(const Wn6Plan& pl, int i) {
    return i == 0 ? 0 : (long long)pl.fb_floats + (long long)(i - 1) * pl.lb_floats;
}
WN_HD int wn6_blob_floats(const Wn6Plan& pl, int i) {
    return i == 0 ? pl.fb_floats : (i < pl.L ? pl.lb_floats : pl.tb_floats);
}
// pair offset of rank r's slice of the exchange published by stage s
WN_HD long long wn6_ex_off(const Wn6Plan& pl, int s, int r) {
    if (s < pl.L) return pl.ex_yx + ((long long)s * pl.CS + r) * pl.rs_yx;
    if (s == pl.L) return pl.ex_sk + (long long)r * pl.rs_sk;
    if (s == pl.L + 1) return pl.ex_h1 + (long long)r * pl.rs_sk;
    return pl.ex_h2 + (long long)r * pl.rs_h2;
}
