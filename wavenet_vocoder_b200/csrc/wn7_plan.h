// wn7_plan.h — execution plan of the synthesis kernel (wn7_kernel.cuh), shared by the host planner/packer and the
// device code.
//
// One generated sample is a chain of L+3 small matrix-vector "stages" with a strict serial dependency (reference
// wavenet.py:296-336).  P thread blocks (one per SM, cooperative launch) each own a fixed slice of the OUTPUT rows
// of every matrix, so a weight never moves between SMs; the stage outputs (a few hundred floats) are exchanged as
// tagged (value, tag) pairs through L2.  What round 2 measured and this plan is shaped by (profiles/r2_*):
//   * an L1-bypassing coherent load costs the issuing warp ~250 cycles and several of them from one lane do not
//     overlap, so the exchange must be read by MANY lanes with ONE 16-byte load each: dedicated polling warps copy
//     the whole stage vector into shared memory and release it with an mbarrier;
//   * a second hop (partial sums through distributed shared memory, the round-2 cluster experiment) costs more than
//     the polling traffic it saves: every block reads the whole vector and evaluates its rows over the FULL K;
//   * cross-warp reductions and block-wide barriers were half of round 1's stage time: a "pass" gives one warp two
//     complete rows (lanes split K, 4 consecutive k per lane and step), so the 32-lane butterfly ends in the lanes
//     that finalise (bias / gate / residual / ReLU) and publish -- no partial sums leave the warp.
// All of it is plain arithmetic on the model shape, so the host tests replay the packed image without a GPU.
#pragma once
#include <stdint.h>

#ifndef WN_HD
#ifdef __CUDACC__
#define WN_HD __host__ __device__ __forceinline__
#else
#define WN_HD inline
#endif
#endif

#define WN7_NCW 8                 // compute warps
#define WN7_MAX_NPW 8             // polling warps (plan.npw of them are launched)
#define WN7_MAX_BT 8
#define WN7_MAX_CI 4              // local-conditioning channels <= 32*WN7_MAX_CI

// stage kinds (which pass list a stage uses)
enum { WN7_K_FIRST = 0, WN7_K_LAYER = 1, WN7_K_TAIL = 2, WN7_K_HEAD1 = 3, WN7_K_HEAD2 = 4, WN7_NKIND = 5 };
// jobs == what the two rows of a pass are and how they are finalised
enum { WN7_J_A0 = 0,   // stage 0: gate pair of layer 0 from x_0                          -> gate, publish y_0
       WN7_J_A = 1,    // gate pair: [M_{s-1} | V_s] x (y_{s-1}, x_{s-1})                 -> gate, publish y_s
       WN7_J_B = 2,    // two residual rows: conv1x1_out_{s-1} x y_{s-1}                  -> (o + b + x_{s-1}) sqrt(.5), publish x_s
       WN7_J_D = 3,    // one older tap of a gate pair of layer s-1 x x_{s-1} (deferred)  -> history ring
       WN7_J_S = 4,    // two skip rows of layer s-1 x y_{s-1} (deferred)                 -> skip accumulator
       WN7_J_SL = 5,   // two skip rows of the last layer (stage L)                       -> total skip, ReLU, publish
       WN7_J_HA = 6,   // two rows of last_conv_layers[1]                                 -> ReLU, publish
       WN7_J_HB = 7,   // two rows of last_conv_layers[3]                                 -> publish
       WN7_NJOB = 8 };

struct Wn7Pass {
    int32_t w_off;        // float offset of the tile inside the stage's blob: [j][row 0..1][lane][4 k]
    int16_t nit;          // k-steps: lane handles k = x_off + 4*(lane + 32*j) .. +3
    int16_t x_off;        // first k of the tile in the stage input vector (multiple of 4)
    int16_t idx;          // owner-local index of the first row of the pass inside its job (pair index for gates)
    int8_t job;           // WN7_J_*
    int8_t deferred;      // 1: off the critical path (J_D, J_S)
};

struct Wn7Plan {
    // ---- model shape (wavenet.py:98-111)
    int L, per_stack, R, G, G2, S, O, kw, C, gin, input_kind, head_kind, Kmix;
    float skip_scale;           // sqrt(1/L), wavenet.py:313
    // ---- grid
    int P, BT, npw;             // blocks, utterances per launch, polling warps
    // ---- rows finalised per block (uniform maxima; the last blocks may own fewer), rounded up to pairs
    int my, mx, ms, mo;         // gate pairs, residual rows, skip (= head-1) rows, head-2 rows
    // ---- stage input vector layout (values; x BT floats in xin, utterance-major: xin[b*xin_vals + k])
    int xoff;                   // offset of the x part behind the y part: round_up(G2, 4)
    int xin_vals;               // values per utterance in one stage-input buffer (multiple of 4, covers every tile)
    // ---- exchange in L2 (pair offsets); vector order == xin order, pair index (k*BT + b)
    int NS;                     // stages per step = L+3
    long long slot_pairs;       // pairs reserved per stage slot
    long long ex_pairs;
    int ex_a, ex_b;             // physical pair index = ((i >> ex_a) << ex_b) + (i & (2^ex_a - 1)): 2^ex_a pairs per
                                // 2^ex_b-pair granule, so that a stage's sectors spread over many L2 slices
    int gate_cycles;            // pollers do not touch a vector earlier than this many cycles after the previous one
    int backoff_ns;             // pause between two polling attempts of a lane
    // ---- pass lists: pass_begin[kind][warp] .. +pass_count, the first pass_crit of them critical
    int npass;
    int pass_begin[WN7_NKIND][WN7_NCW], pass_count[WN7_NKIND][WN7_NCW], pass_crit[WN7_NKIND][WN7_NCW];
    int has_deferred[WN7_NKIND];
    // ---- blobs (floats): first (stage 0), layer (stages 1..L-1), tail (stages L, L+1, L+2)
    int fb_floats, lb_floats, tb_floats, slot_floats;
    long long cta_w_floats;     // packed floats per block = fb + (L-1)*lb + tb
    int nblobs, nres, nring;
    // ---- biases of a block's rows: [L][2my] zb | [L][mx] xb | [L][ms] sb | [ms] | [mo]
    int bo_zb, bo_xb, bo_sb, bo_ha, bo_hb, cta_b_floats;
    // ---- conditioning rows of a block, read from L2 by the conditioning warp: [L][ceil(2my/4)][C][4]
    int qA;
    long long cta_cw_floats;
    // ---- history rings of the older-tap products: one position = 4qA*BT floats
    int ring_in_smem;
    long long ring_pos_total;
    // ---- shared memory map (byte offsets)
    int sm_bar, sm_misc, sm_pass, sm_ringtab, sm_xin, sm_sb, sm_pre, sm_cond, sm_bias, sm_skipacc, sm_xown, sm_hs,
        sm_noise, sm_in, sm_x0w, sm_ring, sm_slots, smem_bytes;
    int nthreads;
};

// balanced split of `rows` over n parts: part p owns [base, base+cnt)
WN_HD void wn7_part(int rows, int n, int p, int& base, int& cnt) {
    const int q = rows / n, r = rows % n;
    base = p * q + (p < r ? p : r);
    cnt = q + (p < r ? 1 : 0);
}
WN_HD int wn7_ceil_div(int a, int b) { return (a + b - 1) / b; }
WN_HD int wn7_dilation(const Wn7Plan& pl, int l) { return 1 << (l % pl.per_stack); }
WN_HD int wn7_kind(const Wn7Plan& pl, int s) {
    return s == 0 ? WN7_K_FIRST : (s < pl.L ? WN7_K_LAYER : (s == pl.L ? WN7_K_TAIL : (s == pl.L + 1 ? WN7_K_HEAD1 : WN7_K_HEAD2)));
}
WN_HD int wn7_blob_of_stage(const Wn7Plan& pl, int s) { return s < pl.L ? s : pl.L; }
WN_HD long long wn7_blob_off(const Wn7Plan& pl, int i) {
    return i == 0 ? 0 : (long long)pl.fb_floats + (long long)(i - 1) * pl.lb_floats;
}
WN_HD int wn7_blob_floats(const Wn7Plan& pl, int i) {
    return i == 0 ? pl.fb_floats : (i < pl.L ? pl.lb_floats : pl.tb_floats);
}
// pair offset of the exchange slot stage s publishes into, and the physical position of pair i inside a slot
WN_HD long long wn7_ex_off(const Wn7Plan& pl, int s) { return (long long)s * pl.slot_pairs; }
WN_HD long long wn7_phys(const Wn7Plan& pl, long long i) {
    return ((i >> pl.ex_a) << pl.ex_b) + (i & ((1LL << pl.ex_a) - 1));
}
// warp roles: [0,npw) pollers, [npw, npw+NCW) compute, then HK (pre-sums / ring positions), TMA, COND
WN_HD int wn7_warp_comp(const Wn7Plan& pl) { return pl.npw; }
WN_HD int wn7_warp_hk(const Wn7Plan& pl) { return pl.npw + WN7_NCW; }
WN_HD int wn7_warp_tma(const Wn7Plan& pl) { return pl.npw + WN7_NCW + 1; }
WN_HD int wn7_warp_cond(const Wn7Plan& pl) { return pl.npw + WN7_NCW + 2; }
