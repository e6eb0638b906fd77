// wn_kernel.cuh — the persistent sm_100a synthesis kernel.
//
// One launch == one WaveNet.incremental_forward() call (reference wavenet.py:215-343): the whole
// T-step loop, including the sampler, runs on the device.  P thread blocks (one per SM,
// cooperative launch) each own a fixed slice of the output rows of every matrix (wn_plan.h).
//
// Per generated sample the blocks run L+3 "stages", each ending in ONE broadcast:
//   stage 0      : x_0 (first 1x1 conv of the fed-back sample, computed by every block) ->
//                  its rows of the current tap of layer 0 -> tanh*sigmoid -> publish y_0
//   stage s<L    : wait for (y_{s-1}, x_{s-1}); then for layer s (modules.py:112-163)
//                    z_s = M_{s-1} y_{s-1} + V_s x_{s-1} + bias + conditioning + queued older taps
//                  where V_s = sqrt(.5) W_s[:,:,kw-1] and M_{s-1} = V_s Wo_{s-1} were folded on the host
//                  (conv1x1_out of the previous layer rides inside the current tap, so the reference's
//                  two dependent GEMVs per layer need one broadcast instead of two), and
//                    x_s = (Wo_{s-1} y_{s-1} + bo + x_{s-1}) sqrt(.5)        (its rows; the residual stream)
//                  -> publish (y_s, x_s) together.
//                  Deferred, off the critical path while the broadcast travels: the OLDER taps'
//                  products W_{s-1}[:,:,k<kw-1] . x_{s-1}(t), queued for steps t+d, t+2d (this replaces
//                  the reference's input shift register, conv.py:32-44, by a queue of OUTPUT partials
//                  private to the block), and its rows of conv1x1_skip_{s-1}, accumulated in layer
//                  order like wavenet.py:312.
//   stage L      : skip rows of the last layer -> total skip * sqrt(1/L) -> ReLU -> publish
//   head 1, 2    : wavenet.py:315-319, one broadcast each
// then the sampler (mixture.py), which every block evaluates redundantly from the same noise so
// no further broadcast is needed.
//
// Inside a block the 8 compute warps are split in two groups that run concurrently:
//   critical group (warps 0-3): wait for the broadcast -> current-tap rows -> gate -> publish.
//       Nothing else sits between two broadcasts.
//   deferred group (warps 4-7): takes (y_{s-1}, x_{s-1}) from a shared-memory stash left by the critical
//       group and does everything that is only needed later (queued older-tap products, skip rows).
// plus one warp that streams weights (TMA) and one that projects the local conditioning.
//
// Exchange protocol: every value travels as an 8-byte (value, tag) pair (tag = step*NE+id+1), written with
// one 8-byte store and polled with 8/16-byte loads, so data and "ready" flag are one atomic word: no
// fences, no separate barrier, one L2 write + one L2 read per hop.  What round 1 measured on B200 and the
// code is shaped by (logs under profiles/):
//   * an L1-bypassing coherent load (SASS LDG.E.64/128.STRONG.GPU) costs the issuing warp ~250 cycles and
//     several of them from one warp do not overlap, however they are scheduled -> the poll time is
//     (loads per thread) x 250 cycles: the vector is read by the 128 threads of the critical group, and with
//     one utterance per launch each thread owns PAIRS of adjacent elements fetched by one 16-byte load
//     (3 loads per thread for config 2);  one-warp-per-quad polling (24 loads per lane), TMA bulk-copy
//     polling and cp.async polling were all measured slower;
//   * a thread that writes several replicas of its value pays ~100-190 cycles per st.relaxed.gpu, and
//     replicas written by different threads still cost more in scattered stores than they save: one copy
//     (`ncopy` = 1; the replica mechanism is kept for experiments);
//   * spreading the pairs over more L2 slices (wn_pair_index) makes no measurable difference.
//
// Weights: fp32, packed per block by the host ("blobs").  A dedicated warp streams the blobs
// into shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a ring
// of slots, running ahead of the compute warps; blobs that fit stay resident for the whole call.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "wn_plan.h"

struct WnPtrs {
    const float* wpack;        // [P][cta_w_floats]
    const float* cwpack;       // [P][cta_cw_floats]
    const float* gbias;        // [B][L][G] = Wg_l . g_b   (NULL without global conditioning)
    const float* first_w;      // scalar input: [R];  one-hot input: transposed [O][R]
    const float* first_b;      // [R]
    uint2* xbuf;               // exchange replicas
    float* ring_g;             // [P][ring floats] when the history rings do not fit in smem
    const int* ringtab;        // [L*(kw-1)*2] : (offset in positions, delay D)
    int* err;                  // [4] device fault word, last tag, block, info
    // ---- per call
    const float* c;
    const float* initial;
    const float* initial_dense;   // one-hot input: (B,O) dense start vector or NULL
    const int* initial_rows;      // one-hot input: (B) start class per utterance or NULL
    const float* test_scalar;
    const int* test_index;
    const float* test_dense;
    const float* u1;
    const float* u2;
    const float* z;
    const float* e;
    float* out_scalar;
    int* out_index;
    float* out_dense;
    float* params_out;
    int B, Btot, b0, T, T_test, initial_index;   // B rows in this launch; noise is strided by Btot
    unsigned flags;
    int noise_kind;
    unsigned long long seed;
    long long timeout_cycles;
    long long* prof;           // optional [P][8] cycle counters (scripts/sweep.py --prof)
    int warp_reverse;          // 1: logical warp = 9 - physical warp (the issue arbiter favours high warp ids)
    int gate_cycles;           // the critical group does not poll an exchange earlier than this after its own publish
    int fast_gate;             // 1: approximate exponentials / division in the gate of the lean stage path
};

#define WN_FLAG_SOFTMAX_ 1u
#define WN_FLAG_QUANTIZE_ 2u

namespace wn {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_pair2(const uint2* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_pair(uint2* p, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag)
                 : "memory");
}
// acquire/release fence at block scope (__threadfence_block() is the sequentially-consistent one: MEMBAR.SC.CTA)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ int ld_flag(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// named barrier over the WN_NT compute threads only (aux warps never join), OR-reducing a flag
__device__ __forceinline__ bool bar_or(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, %2, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"((uint32_t)pred), "n"(WN_NT)
        : "memory");
    return r != 0;
}

// named barrier `ID` over `N` threads, OR-reducing a flag
template <int ID, int N>
__device__ __forceinline__ bool bar_or_n(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.or.pred q, %2, %3, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"((uint32_t)pred), "n"(ID), "n"(N)
        : "memory");
    return r != 0;
}

// The barrier at which the two compute groups of a block meet (id 3, all WN_NT compute threads).  The groups arrive
// from two different loops, which PTX allows (a barrier is its id, not its instruction) but compute-sanitizer's
// synccheck reports as divergence: -DWN_SINGLE_BARRIER_SITE compiles ONE out-of-line instance for that tool
// (profiles/r2_sanitizer_summary.txt).  The shipped build inlines it: the call costs 2 % of a sample (46.5 vs 45.4 us).
#ifdef WN_SINGLE_BARRIER_SITE
__device__ __noinline__ bool bar_groups(bool pred) { return bar_or_n<3, WN_NT>(pred); }
#else
__device__ __forceinline__ bool bar_groups(bool pred) { return bar_or_n<3, WN_NT>(pred); }
#endif

// NA independent value sets reduced in lock step (the shuffles of different sets overlap)
template <int NA, int NV>
__device__ __forceinline__ void reduce_scatter_multi(float (&v)[NA][NV], int lane) {
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
#pragma unroll
                for (int i = 0; i < NV / 2; ++i) {
                    if (i < n) {
                        const float send = hi ? v[a][i] : v[a][i + n];
                        const float keep = hi ? v[a][i + n] : v[a][i];
                        v[a][i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < NA; ++a) v[a][0] += __shfl_xor_sync(0xffffffffu, v[a][0], off);
        }
    }
}

template <int NV>
__device__ __forceinline__ void reduce_scatter(float (&v)[NV], int lane) {
    // butterfly over the 32 lanes; while more than one value is left each step also halves the
    // value set, so NV values cost NV-1+... shuffles instead of 5*NV.  Afterwards v[0] of lane
    // l is the warp-wide sum of value (l >> (5 - log2 NV)).
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < n) {
                    const float send = hi ? v[i] : v[i + n];
                    const float keep = hi ? v[i + n] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// Philox4x32-10 (counter-based; the same (seed, step, utterance, slot) gives the same draw in
// every block, which is what lets all blocks sample redundantly)
__device__ __forceinline__ uint4 philox4(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) {   // (0,1), then mapped like uniform_(1e-5, 1-1e-5)
    const float u = ((r >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return 1e-5f + u * (1.0f - 2e-5f);
}


// slow path of every spin loop: has another block faulted / have we waited too long?
__device__ __noinline__ bool wn_check_abort(volatile int* s_abort, int* err, long long timeout, uint32_t what, int p,
                                            long long& t0) {
    if (*s_abort) return true;
    if (ld_flag(err) != 0) {
        *s_abort = 1;
        return true;
    }
    const long long now = clock64();
    if (t0 == 0) {
        t0 = now;
        return false;
    }
    if (now - t0 > timeout) {
        if (atomicCAS(err, 0, 1) == 0) {
            err[1] = (int)what;
            err[2] = p;
            err[3] = (int)threadIdx.x;
        }
        *s_abort = 1;
        return true;
    }
    return false;
}

// watchdog of the lean poll loops: no clock state on the stack, the elapsed time is bounded from below by the spin count
// (a failed attempt is at least one L2 round trip, > 256 cycles)
__device__ __noinline__ bool wn_poll_check(volatile int* s_abort, int* err, uint32_t spins, long long timeout,
                                           uint32_t what, int p) {
    if (*s_abort) return true;
    if (ld_flag(err) != 0) {
        *s_abort = 1;
        return true;
    }
    if ((long long)spins * 256 > timeout) {
        if (atomicCAS(err, 0, 1) == 0) {
            err[1] = (int)what;
            err[2] = p;
            err[3] = (int)threadIdx.x;
        }
        *s_abort = 1;
        return true;
    }
    return false;
}
// Slow paths of the lean stage code: the common case (barrier already complete / counter already reached) is one
// inline test, everything else lives in these out-of-line loops so that the stage body stays short.
__device__ __noinline__ bool wn_wait_bar_slow(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* err,
                                              long long timeout, uint32_t what, int p) {
    uint32_t spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 255u) == 0 && wn_check_abort(s_abort, err, timeout, what, p, t0)) return false;
    }
    return true;
}
__device__ __noinline__ bool wn_wait_count_slow(volatile int* cnt, int need, unsigned sleep_ns, volatile int* s_abort,
                                                int* err, long long timeout, uint32_t what, int p) {
    uint32_t spins = 0;
    long long t0 = 0;
    while (*cnt < need) {
        if (sleep_ns) __nanosleep(sleep_ns);
        if (((++spins) & 63u) == 0 && wn_check_abort(s_abort, err, timeout, what, p, t0)) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
#define WN_NTC 128                 // threads of one compute group (4 warps)
#define WN_GW 4                    // warps per group
#define WN_DISPATCH_E(EV, ...)                                \
    switch (EV) {                                             \
        case 1: { constexpr int E = 1; __VA_ARGS__ } break;   \
        case 2: { constexpr int E = 2; __VA_ARGS__ } break;   \
        case 4: { constexpr int E = 4; __VA_ARGS__ } break;   \
        default: { constexpr int E = 8; __VA_ARGS__ } break;  \
    }

// LEAN: the lean stage path (crit_loop / def_loop) INSTEAD of the generic one -- a separate, smaller kernel: compiled
// into the same kernel the two paths cost each other registers and instruction-cache footprint (measured: the generic
// path ran 55 % slower with the lean code merely present, profiles/r2_lean_stage_sweeps.txt)
template <int BT, int ER, int EG, bool LEAN = false>
struct Engine {
    static_assert(!LEAN || (BT == 1 && ER % 2 == 0 && EG % 2 == 0), "lean path: one utterance, paired elements");
    static constexpr int NV = 4 * BT;
    static constexpr int NA = (BT >= 8) ? 1 : 2;     // quads reduced together (their shuffle chains overlap)
    const WnPlan& pl;
    const WnPtrs& pp;
    unsigned char* sm;
    int tid, warp, lane, p;
    int gt, gw;                    // thread / warp index inside the compute group
    uint64_t *bar_full, *bar_empty, *bar_cfull, *bar_cempty;
    volatile int* s_abort;
    volatile int* s_stash_cnt;     // stashes published by the critical group (monotonic)
    volatile int* s_ddone_cnt;     // deferred-group warps finished, summed over stages (monotonic)
    int* ringtab;
    float *xs, *ys, *red1, *red2, *sb, *pre, *cond, *skipacc, *hs, *noise, *first, *slots;
    volatile float* ring;
    float* s_in;     // [BT] scalar feedback
    int* s_idx;      // [BT] class feedback
    float* s_dense;  // [BT][O] dense feedback (only without QUANTIZE)
    bool dead;

    __device__ Engine(const WnPlan& pl_, const WnPtrs& pp_, unsigned char* sm_)
        : pl(pl_), pp(pp_), sm(sm_) {
        // logical thread index: the SM's issue arbiter prefers the highest warp id among eligible warps, so the
        // critical group (logical warps 0-3) is mapped onto the highest physical warps when warp_reverse is set
        lane = threadIdx.x & 31;
        warp = pp.warp_reverse ? (WN_NTHREADS / 32 - 1) - (int)(threadIdx.x >> 5) : (int)(threadIdx.x >> 5);
        tid = warp * 32 + lane;
        p = blockIdx.x;
        gt = tid & (WN_NTC - 1);
        gw = warp & (WN_GW - 1);
        const int nslots = pl.nres + pl.nring;
        bar_full = reinterpret_cast<uint64_t*>(sm + pl.sm_bar);
        bar_empty = bar_full + nslots;
        bar_cfull = bar_empty + (pl.nring > 0 ? pl.nring : 1);
        bar_cempty = bar_cfull + 2;
        s_abort = reinterpret_cast<volatile int*>(sm + pl.sm_misc);
        s_stash_cnt = s_abort + 1;
        s_ddone_cnt = s_abort + 2;
        s_in = reinterpret_cast<float*>(sm + pl.sm_in);
        s_idx = reinterpret_cast<int*>(s_in + BT);
        s_dense = reinterpret_cast<float*>(s_idx + BT);
        ringtab = reinterpret_cast<int*>(sm + pl.sm_ringtab);
        xs = reinterpret_cast<float*>(sm + pl.sm_xs);
        ys = xs + 2 * pl.R * BT;
        red1 = reinterpret_cast<float*>(sm + pl.sm_red1);
        red2 = reinterpret_cast<float*>(sm + pl.sm_red2);
        sb = reinterpret_cast<float*>(sm + pl.sm_sb);
        pre = sb + (size_t)pl.L * pl.RA4 * BT;
        cond = reinterpret_cast<float*>(sm + pl.sm_cond);
        skipacc = reinterpret_cast<float*>(sm + pl.sm_skipacc);
        hs = reinterpret_cast<float*>(sm + pl.sm_hs);
        noise = reinterpret_cast<float*>(sm + pl.sm_noise);
        first = reinterpret_cast<float*>(sm + pl.sm_first);
        slots = reinterpret_cast<float*>(sm + pl.sm_slots);
        if (pl.ring_in_smem)
            ring = reinterpret_cast<volatile float*>(sm + pl.sm_ring);
        else
            ring = pp.ring_g + (size_t)p * pl.ring_pos_total * pl.RA4 * BT;
        dead = false;
    }

    // ---- watchdog: a stuck wait sets the device fault word and makes every block unwind
    __device__ __forceinline__ bool check_abort(uint32_t what, long long& t0) {
        return wn_check_abort(s_abort, pp.err, pp.timeout_cycles, what, p, t0);
    }
    // `relaxed` waits (anything off the critical path) back off with nanosleep so that they do not
    // take issue slots and LSU bandwidth from the critical warp of the same SM sub-partition
    template <bool relaxed = false>
    __device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, uint32_t what) {
        uint32_t spins = 0;
        long long t0 = 0;
        while (!mbar_try_wait(bar, parity)) {
            if (relaxed) __nanosleep(64);
            if (((++spins) & (relaxed ? 63u : 255u)) == 0 && check_abort(what, t0)) return false;
        }
        return true;
    }
    // monotonic shared-memory counter (cross-group hand-off inside the block)
    template <bool relaxed = false>
    __device__ __forceinline__ void wait_count(volatile int* cnt, int need, uint32_t what) {
        uint32_t spins = 0;
        long long t0 = 0;
        while (*cnt < need) {
            if (relaxed) __nanosleep(32);
            if (((++spins) & (relaxed ? 63u : 255u)) == 0 && check_abort(what, t0)) {
                dead = true;
                return;
            }
        }
        __threadfence_block();
    }

    // ---- wait for a broadcast vector: thread owns elements k = gt + j*WN_NTC.
    // Which element of a K-vector is slot j of this thread.  With one utterance per launch a thread owns
    // PAIRS of adjacent elements so that one 16-byte load fetches two (value, tag) pairs: every L1-bypassing
    // coherent load costs the issuing warp ~250 cycles and they do not overlap (profiles/r1_v4_*), so the
    // number of loads per thread is what sets the poll time.
    template <int E>
    __device__ __forceinline__ int elem(int j) const {
        if constexpr (BT == 1 && (E % 2) == 0) return 2 * gt + 2 * WN_NTC * (j >> 1) + (j & 1);
        else return gt + j * WN_NTC;
    }
    // `src` is the base of the block's replica, `e0` the first element of the vector
    template <int E>
    __device__ __forceinline__ uint32_t load_vec(const uint2* __restrict__ src, int e0, int K, uint32_t tag,
                                                 float (&x)[E][BT]) {
        uint32_t bad = 0;
        if constexpr (BT == 1 && (E % 2) == 0) {
            uint4 raw[E / 2];
#pragma unroll
            for (int j = 0; j < E / 2; ++j) {
                const int k = elem<E>(2 * j);
                raw[j] = make_uint4(0u, tag, 0u, tag);
                if (k < K) raw[j] = ld_pair2(src + wn_pair_index((long long)(e0 + k)));
            }
#pragma unroll
            for (int j = 0; j < E / 2; ++j) {
                bad |= (raw[j].y ^ tag) | ((elem<E>(2 * j) + 1 < K) ? (raw[j].w ^ tag) : 0u);
                x[2 * j][0] = __uint_as_float(raw[j].x);
                x[2 * j + 1][0] = __uint_as_float(raw[j].z);
            }
        } else {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int k = elem<E>(j);
                if (k < K) {
                    const uint2* s = src + wn_pair_index((long long)(e0 + k) * BT);
                    if constexpr (BT == 1) {
                        const uint2 v = ld_pair(s);
                        x[j][0] = __uint_as_float(v.x);
                        bad |= v.y ^ tag;
                    } else {
#pragma unroll
                        for (int b = 0; b < BT; b += 2) {
                            const uint4 v = ld_pair2(s + b);
                            x[j][b] = __uint_as_float(v.x);
                            bad |= v.y ^ tag;
                            x[j][b + 1] = __uint_as_float(v.z);
                            bad |= v.w ^ tag;
                        }
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < BT; ++b) x[j][b] = 0.f;
                }
            }
        }
        return bad;
    }
    template <int E>
    __device__ __forceinline__ void poll_vec(const uint2* __restrict__ src, int e0, int K, uint32_t tag,
                                             float (&x)[E][BT]) {
        uint32_t spins = 0;
        long long t0 = 0;
        while (load_vec<E>(src, e0, K, tag, x) != 0) {
            if (((++spins) & 63u) == 0 && check_abort(tag, t0)) {
                dead = true;
                return;
            }
        }
    }
    // two vectors of the same exchange (y then x): all loads of an attempt are in flight together
    template <int EA, int EB>
    __device__ __forceinline__ void poll_vec2(const uint2* __restrict__ src, int ea, int KA, float (&a)[EA][BT],
                                              int eb, int KB, float (&b)[EB][BT], uint32_t tag) {
        uint32_t spins = 0;
        long long t0 = 0;
        while (true) {
            const uint32_t bad = load_vec<EA>(src, ea, KA, tag, a) | load_vec<EB>(src, eb, KB, tag, b);
            if (bad == 0) return;
            if (((++spins) & 63u) == 0 && check_abort(tag, t0)) {
                dead = true;
                return;
            }
        }
    }
    template <int E>
    __device__ __forceinline__ void stash(float* dst, int K, const float (&x)[E][BT]) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int k = elem<E>(j);
            if (k < K) {
#pragma unroll
                for (int b = 0; b < BT; ++b) dst[k * BT + b] = x[j][b];
            }
        }
    }
    template <int E>
    __device__ __forceinline__ void unstash(const float* src, int K, float (&x)[E][BT]) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int k = elem<E>(j);
#pragma unroll
            for (int b = 0; b < BT; ++b) x[j][b] = (k < K) ? src[k * BT + b] : 0.f;
        }
    }

    // ---- one row quad (4 rows x K) times the thread's slice of the input vector, accumulated
    template <int E>
    __device__ __forceinline__ void quad_fma(const float* __restrict__ wq /* [K][4] */, int K,
                                             const float (&x)[E][BT], float (&acc)[NV]) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int k = elem<E>(j);
            if (k < K) {
                const float4 w4 = *reinterpret_cast<const float4*>(wq + (size_t)k * 4);
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    acc[0 * BT + b] = fmaf(w4.x, x[j][b], acc[0 * BT + b]);
                    acc[1 * BT + b] = fmaf(w4.y, x[j][b], acc[1 * BT + b]);
                    acc[2 * BT + b] = fmaf(w4.z, x[j][b], acc[2 * BT + b]);
                    acc[3 * BT + b] = fmaf(w4.w, x[j][b], acc[3 * BT + b]);
                }
            }
        }
    }
    // after the warp reduction lane (v << (5-M)) holds value v of the quad; warp gw's partial goes
    // to red[(q*NV + v)*4 + gw]
    __device__ __forceinline__ void quad_store(const float (&acc)[NV], int q, float* __restrict__ red) {
        constexpr int M = ilog2c(NV);
        if ((lane & ((32 >> M) - 1)) == 0) red[(q * NV + (lane >> (5 - M))) * WN_GW + gw] = acc[0];
    }
    __device__ __forceinline__ float red_sum(const float* red, int rowidx, int b) const {
        const int v = (rowidx >> 2) * NV + (rowidx & 3) * BT + b;
        const float4 a = *reinterpret_cast<const float4*>(red + v * WN_GW);
        return (a.x + a.y) + (a.z + a.w);
    }
    // simple GEMV over NQ quads, two quads per pass so that their shuffle chains overlap
    template <int E>
    __device__ __forceinline__ void gemv(const float* __restrict__ w, int NQ, int K, const float (&x)[E][BT],
                                         float* __restrict__ red, int q0 = 0) {
        for (int q = 0; q < NQ; q += NA) {
            float acc[NA][NV];
#pragma unroll
            for (int h = 0; h < NA; ++h) {
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[h][v] = 0.f;
                if (q + h < NQ) quad_fma<E>(w + (size_t)(q + h) * K * 4, K, x, acc[h]);
            }
            reduce_scatter_multi<NA, NV>(acc, lane);
#pragma unroll
            for (int h = 0; h < NA; ++h)
                if (q + h < NQ) quad_store(acc[h], q0 + q + h, red);
        }
    }
    __device__ __forceinline__ void publish(int elem, int b, int copy, float v, uint32_t tag) {
        st_pair(pp.xbuf + (size_t)copy * pl.copy_stride_pairs + wn_pair_index((long long)elem * BT + b), v, tag);
    }

    // ---- weight slots.  Every compute warp walks the blobs 0..L of every step in order; the ones that
    // stream go through the ring, tracked by a running (slot, parity) pair (no divisions).
    int rs_slot = 0;
    uint32_t rs_par = 0;
    __device__ __forceinline__ const float* acquire_blob(int t, int i) {
        if (i < pl.nres) {
            if (t == 0 && !wait_bar(&bar_full[i], 0, 0x80000000u | (uint32_t)i)) dead = true;
            return slots + (size_t)i * pl.slot_floats;
        }
        const int slot = pl.nres + rs_slot;
        if (!wait_bar(&bar_full[slot], rs_par, 0x80000000u | (uint32_t)i)) dead = true;
        return slots + (size_t)slot * pl.slot_floats;
    }
    __device__ __forceinline__ void release_blob(int t, int i) {
        if (i >= pl.nres) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[rs_slot]);
            if (++rs_slot == pl.nring) {
                rs_slot = 0;
                rs_par ^= 1u;
            }
        }
    }

    // lean variants (inline fast path, out-of-line spin)
    __device__ __forceinline__ const float* acquire_lean(int t, int i) {
        uint64_t* bar;
        uint32_t par;
        const float* ptr;
        if (i < pl.nres) {
            ptr = slots + (size_t)i * pl.slot_floats;
            if (t != 0) return ptr;
            bar = &bar_full[i];
            par = 0;
        } else {
            const int slot = pl.nres + rs_slot;
            ptr = slots + (size_t)slot * pl.slot_floats;
            bar = &bar_full[slot];
            par = rs_par;
        }
        if (!mbar_try_wait(bar, par)) {
            if (!wn_wait_bar_slow(bar, par, s_abort, pp.err, pp.timeout_cycles, 0x80000000u | (uint32_t)i, p)) dead = true;
        }
        return ptr;
    }
    __device__ __forceinline__ void count_lean(volatile int* cnt, int need, unsigned sleep_ns, uint32_t what) {
        if (*cnt < need) {
            if (!wn_wait_count_slow(cnt, need, sleep_ns, s_abort, pp.err, pp.timeout_cycles, what, p)) dead = true;
        }
        fence_cta();
    }

    // ======================================================================================
    // weight streaming warp
    // ======================================================================================
    __device__ void tma_loop() {
        if (lane != 0) return;
        const float* base = pp.wpack + (size_t)p * pl.cta_w_floats;
        for (int i = 0; i < pl.nres; ++i) {
            const uint32_t bytes = (uint32_t)wn_blob_floats(pl, i) * 4u;
            mbar_expect_tx(&bar_full[i], bytes);
            bulk_g2s(slots + (size_t)i * pl.slot_floats, base + wn_blob_off(pl, i), bytes, &bar_full[i]);
        }
        const int nstream = pl.nblobs - pl.nres;
        if (nstream <= 0) return;
        const uint32_t total = (uint32_t)pp.T * (uint32_t)nstream;
        int i = pl.nres;
        for (uint32_t js = 0; js < total; ++js) {
            const uint32_t s = js % (uint32_t)pl.nring, u = js / (uint32_t)pl.nring;
            if (u > 0) {
                if (!wait_bar<true>(&bar_empty[s], (u - 1) & 1u, 0x40000000u | s)) return;
            }
            const uint32_t bytes = (uint32_t)wn_blob_floats(pl, i) * 4u;
            uint64_t* fb = &bar_full[pl.nres + s];
            mbar_expect_tx(fb, bytes);
            bulk_g2s(slots + (size_t)(pl.nres + s) * pl.slot_floats, base + wn_blob_off(pl, i), bytes, fb);
            if (++i == pl.nblobs) i = pl.nres;
        }
    }

    // ======================================================================================
    // conditioning warp: cond[t&1][l][row][b] = Wc_l[rows] . c_t  (modules.py:141-145), one step
    // ahead of the compute warps; weights come straight from L2 (they are read once per step)
    // ======================================================================================
    __device__ void cond_loop() {
        const int C = pl.C, L = pl.L, T = pp.T, B = pp.B;
        constexpr int NV = 4 * BT;
        constexpr int M = ilog2c(NV);
        const float* cw = pp.cwpack + (size_t)p * pl.cta_cw_floats;
        for (int t = 0; t < T; ++t) {
            const int par = t & 1, u = t >> 1;
            if (u > 0) {
                if (!wait_bar<true>(&bar_cempty[par], (u - 1) & 1u, 0x20000000u)) return;
            }
            float ct[BT][WN_MAX_CI];
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int i = 0; i < WN_MAX_CI; ++i) {
                    const int ch = lane + 32 * i;
                    ct[b][i] = (b < B && ch < C) ? __ldg(pp.c + ((size_t)b * T + t) * C + ch) : 0.f;
                }
            float* dst = cond + (size_t)par * L * pl.RA4 * BT;
            for (int l = 0; l < L; ++l) {
                for (int q = 0; q < pl.NQ_A; ++q) {
                    float acc[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
                    const float* wq = cw + ((size_t)(l * pl.NQ_A + q) * C) * 4;
#pragma unroll
                    for (int i = 0; i < WN_MAX_CI; ++i) {
                        const int ch = lane + 32 * i;
                        if (ch < C) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wq + (size_t)ch * 4));
#pragma unroll
                            for (int b = 0; b < BT; ++b) {
                                acc[0 * BT + b] = fmaf(w4.x, ct[b][i], acc[0 * BT + b]);
                                acc[1 * BT + b] = fmaf(w4.y, ct[b][i], acc[1 * BT + b]);
                                acc[2 * BT + b] = fmaf(w4.z, ct[b][i], acc[2 * BT + b]);
                                acc[3 * BT + b] = fmaf(w4.w, ct[b][i], acc[3 * BT + b]);
                            }
                        }
                    }
                    reduce_scatter<NV>(acc, lane);
                    if ((lane & ((32 >> M) - 1)) == 0) {
                        const int v = lane >> (5 - M);   // = row_in_quad*BT + b
                        dst[((size_t)l * pl.RA4 + q * 4 + v / BT) * BT + (v % BT)] = acc[0];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_cfull[par]);
        }
    }

    // ======================================================================================
    // sampler (one warp per utterance; every block computes the same thing)
    // ======================================================================================
    __device__ __forceinline__ void warp_argmax(float& best, int& bi) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
    }
    // noise for step t of utterance b into noise[b][*]; layout [u1(0..K-1) | u2 or z] or [e(0..O-1)]
    __device__ void fetch_noise(int t, int b) {
        float* nz = noise + (size_t)b * (pl.O + 2);
        const int B = pp.Btot, K = pl.Kmix, O = pl.O;
        const uint32_t ub = (uint32_t)(pp.b0 + b);
        const bool replay = pp.noise_kind == 0;
        const uint2 key = make_uint2((uint32_t)pp.seed, (uint32_t)(pp.seed >> 32));
        if (b >= pp.B) {   // padding row of the batch tile: harmless constants
            for (int i = lane; i < O + 2; i += 32) nz[i] = 0.5f;
            return;
        }
        if (pl.head_kind == 2) {
            for (int i = lane; i < O; i += 32) {
                float e;
                if (replay) e = pp.e ? __ldg(pp.e + ((size_t)t * B + b) * O + i) : 1.0f;
                else {
                    const uint4 r = philox4(make_uint4((uint32_t)t, ub, (uint32_t)i, 2u), key);
                    e = -logf(u01(r.x));
                }
                nz[i] = e;
            }
            return;
        }
        const bool mix = (pl.head_kind == 0) || (K > 1);
        if (mix) {
            for (int i = lane; i < K; i += 32) {
                float u;
                if (replay) u = __ldg(pp.u1 + ((size_t)t * B + b) * K + i);
                else u = u01(philox4(make_uint4((uint32_t)t, ub, (uint32_t)i, 0u), key).x);
                nz[i] = u;
            }
        }
        if (lane == 0) {
            float v;
            if (pl.head_kind == 0) {
                if (replay) v = __ldg(pp.u2 + (size_t)t * B + b);
                else v = u01(philox4(make_uint4((uint32_t)t, ub, 0u, 1u), key).x);
            } else {
                if (replay) v = __ldg(pp.z + (size_t)t * B + b);
                else {
                    const uint4 r = philox4(make_uint4((uint32_t)t, ub, 0u, 1u), key);
                    v = sqrtf(-2.f * logf(u01(r.x))) * cospif(2.f * u01(r.y));   // Box-Muller
                }
            }
            nz[K] = v;
        }
    }
    // draw sample of utterance b from hs[:, b]; sets the feedback for step t+1 and writes outputs
    __device__ void sample_utt(int t, int b) {
        const int O = pl.O, K = pl.Kmix, T = pp.T;
        const float* nz = noise + (size_t)b * (pl.O + 2);
        const bool writer = (p == 0);
        if (pl.head_kind == 2) {
            const bool softmax = (pp.flags & WN_FLAG_SOFTMAX_) != 0, quant = (pp.flags & WN_FLAG_QUANTIZE_) != 0;
            // F.softmax (wavenet.py:332): exp(h - max) / sum
            if (softmax) {
                float m = -INFINITY;
                for (int i = lane; i < O; i += 32) m = fmaxf(m, hs[i * BT + b]);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                float s = 0.f;
                for (int i = lane; i < O; i += 32) {
                    const float e = expf(hs[i * BT + b] - m);
                    hs[i * BT + b] = e;
                    s += e;
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                for (int i = lane; i < O; i += 32) hs[i * BT + b] = hs[i * BT + b] / s;
            }
            if (quant) {
                // OneHotCategorical(p).sample() (wavenet.py:334-335): renormalise, argmax(p / Exp(1))
                float sp = 0.f;
                for (int i = lane; i < O; i += 32) sp += hs[i * BT + b];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, off);
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int i = lane; i < O; i += 32) {
                    const float r = (hs[i * BT + b] / sp) / nz[i];
                    if (r > best) {
                        best = r;
                        bi = i;
                    }
                }
                warp_argmax(best, bi);
                if (bi >= O) bi = 0;
                if (lane == 0) {
                    if (writer && b < pp.B) pp.out_index[(size_t)b * T + t] = bi;
                    s_idx[b] = (t + 1 < pp.T_test && b < pp.B) ? (pp.test_index ? pp.test_index[(size_t)b * pp.T_test + t + 1] : -1)
                                                               : bi;
                }
            } else {
                for (int i = lane; i < O; i += 32) {
                    const float v = hs[i * BT + b];
                    if (writer && b < pp.B) pp.out_dense[((size_t)b * O + i) * T + t] = v;
                    s_dense[b * O + i] = v;
                }
                if (lane == 0)
                    s_idx[b] = (t + 1 < pp.T_test && b < pp.B && pp.test_index)
                                   ? pp.test_index[(size_t)b * pp.T_test + t + 1] : -1;
            }
            // teacher forcing with dense rows overrides the feedback
            if (t + 1 < pp.T_test && pp.test_dense != nullptr && b < pp.B) {
                for (int i = lane; i < O; i += 32)
                    s_dense[b * O + i] = pp.test_dense[((size_t)b * pp.T_test + t + 1) * O + i];
                if (lane == 0) s_idx[b] = -1;
            }
            return;
        }
        // ---- scalar heads
        float mean, ls;
        const bool mix = (pl.head_kind == 0) || (K > 1);
        if (mix) {
            // Gumbel-max over the K mixture logits (mixture.py:138-140 / :247-249)
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = lane; i < K; i += 32) {
                const float g = hs[i * BT + b] - logf(-logf(nz[i]));
                if (g > best) {
                    best = g;
                    bi = i;
                }
            }
            warp_argmax(best, bi);
            if (bi >= K) bi = 0;
            mean = hs[(K + bi) * BT + b];        // mixture.py:143-146 one-hot select
            ls = hs[(2 * K + bi) * BT + b];
        } else if (O == 2) {
            mean = hs[0 * BT + b];               // mixture.py:258-259
            ls = hs[1 * BT + b];
        } else {
            mean = hs[1 * BT + b];               // mixture.py:260-261 (C == 3)
            ls = hs[2 * BT + b];
        }
        float xv;
        if (pl.head_kind == 0) {
            const float u = nz[K];
            // mixture.py:152  x = mu + exp(s) * (log u - log(1-u)); separate roundings as in torch
            xv = __fadd_rn(mean, __fmul_rn(expf(ls), __fsub_rn(logf(u), logf(__fsub_rn(1.0f, u)))));
        } else {
            // mixture.py:265-267  Normal(mu, exp(s)).sample() == z * sigma + mu
            xv = __fadd_rn(__fmul_rn(nz[K], expf(ls)), mean);
        }
        xv = fminf(fmaxf(xv, -1.0f), 1.0f);      // mixture.py:154 / :269
        if (lane == 0) {
            if (writer && b < pp.B) pp.out_scalar[(size_t)b * T + t] = xv;
            s_in[b] = (t + 1 < pp.T_test && b < pp.B) ? pp.test_scalar[(size_t)b * pp.T_test + t + 1] : xv;
        }
    }

    // ======================================================================================
    // compute groups
    // ======================================================================================
    __device__ __forceinline__ static int efor(int K) {
        return K <= WN_NTC ? 1 : (K <= 2 * WN_NTC ? 2 : (K <= 4 * WN_NTC ? 4 : 8));
    }

    // x_0 = first 1x1 conv of the fed-back sample (wavenet.py:308); every block computes all of it
    __device__ __forceinline__ void make_x0(float (&x)[ER][BT]) {
        const int R = pl.R, O = pl.O;
#pragma unroll
        for (int j = 0; j < ER; ++j) {
            const int k = elem<ER>(j);
#pragma unroll
            for (int b = 0; b < BT; ++b) x[j][b] = 0.f;
            if (k < R) {
                if (pl.input_kind == 0) {
#pragma unroll
                    for (int b = 0; b < BT; ++b) x[j][b] = fmaf(first[k], s_in[b], first[R + k]);
                } else {
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const int idx = min(s_idx[b], O - 1);      // class ids are range-checked on the host where it can
                        if (idx >= 0) {
                            // one-hot input: the GEMV is a column gather
                            x[j][b] = __ldg(pp.first_w + (size_t)idx * R + k) + first[R + k];
                        } else {
                            float a = 0.f;
                            for (int o = 0; o < O; ++o)
                                a = fmaf(__ldg(pp.first_w + (size_t)o * R + k), s_dense[b * O + o], a);
                            x[j][b] = a + first[R + k];
                        }
                    }
                }
            }
        }
    }
    // modules.py:154  tanh(a) * sigmoid(g) with a single division:
    //   (1 - e^{-2a}) / ((1 + e^{-2a}) (1 + e^{-g}));  |a| is clamped where tanh has saturated in fp32.
    // Absolute error ~1e-7 (the subtraction 1 - e^{-2a} loses relative, not absolute, accuracy near 0).
    // ex2.approx / rcp.approx version (~1e-6 absolute error), selected at run time by pp.fast_gate
    __device__ __forceinline__ static float gate_fast(float a, float g) {
        const float ac = fminf(fmaxf(a, -15.0f), 15.0f);
        const float ea = __expf(-2.0f * ac), eg = __expf(-g);
        return __fdividef(1.0f - ea, (1.0f + ea) * (1.0f + eg));
    }
    __device__ __forceinline__ static float gate(float a, float g) {
        const float ac = fminf(fmaxf(a, -15.0f), 15.0f);
#ifdef WN_FAST_GATE
        // ex2.approx based exponentials and an approximate division: ~2 ulp each, well inside the 1e-6
        // absolute error the parity tests observe
        const float ea = __expf(-2.0f * ac), eg = __expf(-g);
        return __fdividef(1.0f - ea, (1.0f + ea) * (1.0f + eg));
#else
        const float ea = expf(-2.0f * ac), eg = expf(-g);
        return (1.0f - ea) / ((1.0f + ea) * (1.0f + eg));
#endif
    }
    // Everything of z_l(t) that does not depend on step t's broadcasts: (folded) bias + global conditioning
    // + local conditioning projection + the queued products of the older taps.  The deferred group builds
    // the whole table for step `t` while the critical group is still in the head of step t-1.
    __device__ void build_pre(int t) {
        const int L = pl.L, RA4 = pl.RA4, kw = pl.kw, n = L * pl.RA * BT;
        if (pl.C > 0) {
            if (!wait_bar<true>(&bar_cfull[t & 1], (uint32_t)(t >> 1) & 1u, 0x10000000u)) dead = true;
        }
        const float* cd = cond + (size_t)(t & 1) * L * RA4 * BT;
        for (int i = gt; i < n; i += WN_NTC) {
            const int b = i % BT, rr = (i / BT) % pl.RA, l = i / (BT * pl.RA);
            const int idx = (l * RA4 + rr) * BT + b;
            float v = sb[idx];
            if (pl.C > 0) v += cd[idx];
            for (int k = 0; k < kw - 1; ++k) {
                const int e = (l * (kw - 1) + k) * 3;
                int pos = ringtab[e + 2];
                if (t > 0) { ++pos; if (pos == ringtab[e + 1]) pos = 0; }     // the table still holds (t-1) mod delay
                v += ring[((size_t)ringtab[e] + pos) * RA4 * BT + rr * BT + b];
            }
            pre[idx] = v;
        }
        __syncwarp();
        if (pl.C > 0) {
            // the conditioning warp may refill cond[t&1] once all four deferred warps are done with it
            if (lane == 0) mbar_arrive(&bar_cempty[t & 1]);
        }
    }

    // --------------------------------------------------------------------------------------
    // critical group (warps 0-3)
    // --------------------------------------------------------------------------------------
    __device__ void crit_loop() {
        const int L = pl.L, R = pl.R, G2 = pl.G2, S = pl.S, O = pl.O, T = pp.T, P = pl.P, ncopy = pl.ncopy;
        int y0, ny, x0r, nx, s0, ns, a0, na, b0, nb;
        wn_part(G2, P, p, y0, ny);
        wn_part(R, P, p, x0r, nx);
        wn_part(S, P, p, s0, ns);
        wn_part(S, P, p, a0, na);
        wn_part(O, P, p, b0, nb);
        const int ES = efor(S), EO = efor(O);
        const uint2* xin = pp.xbuf + (size_t)(p % ncopy) * pl.copy_stride_pairs;
        const uint32_t NEID = (uint32_t)L + 3u;
        const int YX = G2 + R;
        const float RSQRT2 = 0.70710678118654752440f;         // math.sqrt(0.5), modules.py:162
        const int NQ_A = pl.NQ_A, NQ_BO = pl.NQ_BO, nqc = NQ_A + NQ_BO;
        // finalizer roles: threads [0,64) publish gate outputs (and skip / head rows), [64,128) the
        // residual rows; local index u -> (item, replica)
        const int fl = gt & 63;
        const bool grpA = gt < 64;
        auto role = [&](int nitems, int& item, int& copy) {
            item = -1; copy = 0;
            if (nitems > 0 && fl < nitems * ncopy) { item = fl % nitems; copy = fl / nitems; }
        };
        int it_y, cp_y, it_x, cp_x, it_s, cp_s, it_a, cp_a, it_b, cp_b;
        role(ny * BT, it_y, cp_y);
        role(nx * BT, it_x, cp_x);
        role(ns * BT, it_s, cp_s);
        role(na * BT, it_a, cp_a);
        role(nb * BT, it_b, cp_b);
        if (!grpA) it_y = it_s = it_a = it_b = -1;
        if (grpA) it_x = -1;
        const int pa_idx = it_y >= 0 ? (2 * (it_y / BT)) * BT + (it_y % BT) : 0;   // [row a_j][b]; row b_j is BT further
        float xr[ER][BT], yr[EG][BT];
        const bool prof = (pp.prof != nullptr) && tid == 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = 0;
#define WN_TICK(i) if (prof) { const long long now_ = clock64(); pc[i] += now_ - tc; tc = now_; }
        int nstash = 0;      // stashes published so far == deferred stages started
        int ndone = 0;       // deferred stages this group has waited for
        long long t_pub = clock64();
        if (bar_groups(false)) return;      // the deferred group has built the pre-sums of step 0

        // ---- lean stage path (pl.lean, set by wn_host.cu: one utterance, exact vector lengths, one gate quad and one
        // residual quad per block).  A stage is a chain of dependent instructions executed by one warp per SM
        // sub-partition, ~5 cycles each (profiles/r2_v5_instruction_profile.txt), so what a stage costs beyond the
        // exchange is its LENGTH IN INSTRUCTIONS: same arithmetic and order as the generic loop below, with every
        // address that does not depend on the stage computed once here, no per-element bounds checks, and the spin
        // loops out of line.
        constexpr bool LEAN_T = LEAN;
        constexpr bool lean = LEAN;
        constexpr int LY = LEAN_T ? EG / 2 : 1, LX = LEAN_T ? ER / 2 : 1;
        long long lk_y[LY], lk_x[LX], l_inc = 0;
        const uint2* l_in0 = xin;
        uint2 *l_puby = pp.xbuf, *l_pubx = pp.xbuf;
        const float* l_pre = pre;
        const uint2 *l_sk0 = xin, *l_h10 = xin;
        uint2 *l_pubs = pp.xbuf, *l_puba = pp.xbuf, *l_pubb = pp.xbuf;
        int l_zy = 0, l_zx = 0, l_xo = 0, l_redw = 0, l_reda = 0, l_redx = 0, l_xb = 0, l_xs = 0;
#pragma unroll
        for (int j = 0; j < LY; ++j) lk_y[j] = 0;
#pragma unroll
        for (int j = 0; j < LX; ++j) lk_x[j] = 0;
        if (lean) {
#pragma unroll
            for (int j = 0; j < LY; ++j) lk_y[j] = wn_pair_index((long long)(2 * gt + 2 * WN_NTC * j));
#pragma unroll
            for (int j = 0; j < LX; ++j) lk_x[j] = wn_pair_index((long long)(G2 + 2 * gt + 2 * WN_NTC * j));
            l_inc = wn_pair_index((long long)YX);
            l_in0 = xin + wn_pair_index((long long)pl.ex_yx);
            l_puby = pp.xbuf + wn_pair_index((long long)pl.ex_yx) + wn_pair_index((long long)(y0 + max(it_y, 0)));
            l_pubx = pp.xbuf + wn_pair_index((long long)pl.ex_yx) + wn_pair_index((long long)(G2 + x0r + max(it_x, 0)));
            l_pre = pre + 2 * max(it_y, 0);
            l_zy = pl.lb_Zy + 8 * gt;
            l_zx = pl.lb_Zx + 8 * gt;
            l_xo = pl.lb_Xo + 8 * gt;
            l_redw = (lane >> 2) * WN_GW + gw;
            l_reda = 2 * max(it_y, 0) * WN_GW;
            l_redx = (4 + max(it_x, 0)) * WN_GW;
            l_xb = pl.lb_xb + max(it_x, 0);
            l_xs = x0r + max(it_x, 0);
            l_sk0 = xin + wn_pair_index((long long)pl.ex_sk);
            l_h10 = xin + wn_pair_index((long long)pl.ex_h1);
            l_pubs = pp.xbuf + wn_pair_index((long long)(pl.ex_sk + s0 + max(it_s, 0)));
            l_puba = pp.xbuf + wn_pair_index((long long)(pl.ex_h1 + a0 + max(it_a, 0)));
            l_pubb = pp.xbuf + wn_pair_index((long long)(pl.ex_h2 + b0 + max(it_b, 0)));
        }
        // lean polls: the thread's pairs of the y part / the x part of one exchange, straight into registers.  A load
        // whose two tags were good is not issued again (every attempt of every block lands on the same few L2 lines,
        // and those reads are what the publishing stores queue behind)
        auto lean_poll_y = [&](const uint2* base, uint32_t tag) {
            uint4 ry[LY];
            uint32_t need = (1u << LY) - 1u, spins = 0;
            while (true) {
#pragma unroll
                for (int j = 0; j < LY; ++j)
                    if (need & (1u << j)) ry[j] = ld_pair2(base + lk_y[j]);
#pragma unroll
                for (int j = 0; j < LY; ++j)
                    if (((ry[j].y ^ tag) | (ry[j].w ^ tag)) == 0u) need &= ~(1u << j);
                if (need == 0u) break;
                if (((++spins) & 63u) == 0 && wn_poll_check(s_abort, pp.err, spins, pp.timeout_cycles, tag, p)) {
                    dead = true;
                    break;
                }
            }
#pragma unroll
            for (int j = 0; j < LY; ++j) {
                yr[2 * j][0] = __uint_as_float(ry[j].x);
                yr[2 * j + 1][0] = __uint_as_float(ry[j].z);
            }
        };
        auto lean_poll_x = [&](const uint2* base, uint32_t tag) {
            uint4 rx[LX];
            uint32_t need = (1u << LX) - 1u, spins = 0;
            while (true) {
#pragma unroll
                for (int j = 0; j < LX; ++j)
                    if (need & (1u << j)) rx[j] = ld_pair2(base + lk_x[j]);
#pragma unroll
                for (int j = 0; j < LX; ++j)
                    if (((rx[j].y ^ tag) | (rx[j].w ^ tag)) == 0u) need &= ~(1u << j);
                if (need == 0u) break;
                if (((++spins) & 63u) == 0 && wn_poll_check(s_abort, pp.err, spins, pp.timeout_cycles, tag | 0x40000000u, p)) {
                    dead = true;
                    break;
                }
            }
#pragma unroll
            for (int j = 0; j < LX; ++j) {
                xr[2 * j][0] = __uint_as_float(rx[j].x);
                xr[2 * j + 1][0] = __uint_as_float(rx[j].z);
            }
        };
        // one row quad times the y-shaped / x-shaped register vector (wq already offset by the thread's 8*gt floats)
        auto lean_quad_y = [&](const float* wq, float (&a4)[4]) {
#pragma unroll
            for (int j = 0; j < 2 * LY; ++j) {
                const float4 a = *reinterpret_cast<const float4*>(wq + (2 * WN_NTC * (j >> 1) + (j & 1)) * 4);
                const float v = yr[j][0];
                a4[0] = fmaf(a.x, v, a4[0]); a4[1] = fmaf(a.y, v, a4[1]);
                a4[2] = fmaf(a.z, v, a4[2]); a4[3] = fmaf(a.w, v, a4[3]);
            }
        };
        auto lean_quad_x = [&](const float* wq, float (&a4)[4]) {
#pragma unroll
            for (int j = 0; j < 2 * LX; ++j) {
                const float4 a = *reinterpret_cast<const float4*>(wq + (2 * WN_NTC * (j >> 1) + (j & 1)) * 4);
                const float v = xr[j][0];
                a4[0] = fmaf(a.x, v, a4[0]); a4[1] = fmaf(a.y, v, a4[1]);
                a4[2] = fmaf(a.z, v, a4[2]); a4[3] = fmaf(a.w, v, a4[3]);
            }
        };
        auto quad_sum = [](const float* r) {
            const float4 q = *reinterpret_cast<const float4*>(r);
            return (q.x + q.y) + (q.z + q.w);
        };

        for (int t = 0; t < T; ++t) {
            const uint32_t tagbase = (uint32_t)t * NEID + 1u;
            if (prof) tc = clock64();
            bool step_dead = false;
            do {
                make_x0(xr);
                WN_TICK(7);
                // ------------------------------------------------------------ stage 0: layer 0 from x_0
                if constexpr (LEAN_T) {
                    {
                        const float* W = acquire_lean(t, 0);
                        const float pre_a = l_pre[0], pre_b = l_pre[1];
                        float a4[4] = {0.f, 0.f, 0.f, 0.f};
                        lean_quad_x(W + pl.fb_Zx + 8 * gt, a4);
                        reduce_scatter<4>(a4, lane);
                        if ((lane & 7) == 0) red1[(lane >> 3) * WN_GW + gw] = a4[0];
                        WN_TICK(1);
                        if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                        WN_TICK(2);
                        if (it_y >= 0) {
                            const float a = quad_sum(red1 + l_reda) + pre_a;
                            const float g = quad_sum(red1 + l_reda + WN_GW) + pre_b;
                            st_pair(l_puby, pp.fast_gate ? gate_fast(a, g) : gate(a, g), tagbase);
                        }
                        release_blob(t, 0);
                        WN_TICK(3);
                    }
                } else {
                    const float* W = acquire_blob(t, 0);
                    float pre_a = 0.f, pre_b = 0.f;
                    if (it_y >= 0) { pre_a = pre[pa_idx]; pre_b = pre[pa_idx + BT]; }
                    float* r1 = red1;                                   // partials buffers alternate by stage
                    gemv<ER>(W + pl.fb_Zx, NQ_A, R, xr, r1);
                    WN_TICK(1);
                    if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                    WN_TICK(2);
                    if (it_y >= 0) {
                        const int fr = it_y / BT, fb = it_y % BT;
                        const float a = red_sum(r1, 2 * fr, fb) + pre_a;
                        const float g = red_sum(r1, 2 * fr + 1, fb) + pre_b;
                        publish(pl.ex_yx + y0 + fr, fb, cp_y, gate(a, g), tagbase + wn_eid_yx(0));
                    }
                    release_blob(t, 0);
                    t_pub = clock64();
                    WN_TICK(3);
                }
                // ------------------------------------------------------------ stages 1..L-1, lean path
                int s_next = 1;
                if constexpr (LEAN_T) {
                    if (lean) {
                        s_next = L;
                        for (int s = 1; s < L; ++s) {
                            const float* W = acquire_lean(t, s);
                            const float pre_a = l_pre[s * pl.RA4], pre_b = l_pre[s * pl.RA4 + 1];
                            const uint32_t tag = tagbase + (uint32_t)(s - 1);
                            const uint2* base = l_in0 + (long long)(s - 1) * l_inc;
                            float* xst = xs + (s & 1) * R;
                            float* yst = ys + (s & 1) * G2;
                            // x_{s-1} was published by the DEFERRED groups half a stage ago (def_loop): normally one attempt,
                            // and its product is formed while y_{s-1} is still travelling (x_0 is already in registers)
                            if (s >= 2) lean_poll_x(base, tag);
#pragma unroll
                            for (int j = 0; j < LX; ++j)
                                *reinterpret_cast<float2*>(xst + 2 * gt + 2 * WN_NTC * j) = make_float2(xr[2 * j][0], xr[2 * j + 1][0]);
                            float a4[4] = {0.f, 0.f, 0.f, 0.f};
                            lean_quad_x(W + l_zx, a4);
                            float4 wy[2 * LY];          // weights of the y part: in registers before the data arrives
#pragma unroll
                            for (int j = 0; j < 2 * LY; ++j)
                                wy[j] = *reinterpret_cast<const float4*>(W + l_zy + (2 * WN_NTC * (j >> 1) + (j & 1)) * 4);
                            if (pp.gate_cycles > 0) { while (clock64() - t_pub < pp.gate_cycles) {} }
                            WN_TICK(4);
                            lean_poll_y(base, tag);
                            WN_TICK(0);
#pragma unroll
                            for (int j = 0; j < 2 * LY; ++j) {
                                const float yv = yr[j][0];
                                a4[0] = fmaf(wy[j].x, yv, a4[0]); a4[1] = fmaf(wy[j].y, yv, a4[1]);
                                a4[2] = fmaf(wy[j].z, yv, a4[2]); a4[3] = fmaf(wy[j].w, yv, a4[3]);
                            }
                            reduce_scatter<4>(a4, lane);
                            float* r1 = red1 + (s & 1) * pl.red1_floats;
                            if ((lane & 7) == 0) r1[(lane >> 3) * WN_GW + gw] = a4[0];
#pragma unroll
                            for (int j = 0; j < LY; ++j)
                                *reinterpret_cast<float2*>(yst + 2 * gt + 2 * WN_NTC * j) = make_float2(yr[2 * j][0], yr[2 * j + 1][0]);
                            if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                            if (gt == 0) {
                                fence_cta();
                                *s_stash_cnt = nstash + 1;
                            }
                            ++nstash;
                            WN_TICK(1);
                            if (it_y >= 0) {
                                const float a = quad_sum(r1 + l_reda) + pre_a;
                                const float g = quad_sum(r1 + l_reda + WN_GW) + pre_b;
                                st_pair(l_puby + (long long)s * l_inc, pp.fast_gate ? gate_fast(a, g) : gate(a, g), tag + 1u);
                            }
                            release_blob(t, s);
                            if (pp.gate_cycles > 0) t_pub = clock64();
                            if (s >= 2) { count_lean(s_ddone_cnt, WN_GW * (ndone + 1), 0u, 0x08000000u); ++ndone; }
                            WN_TICK(3);
                        }
                    }
                }
                // ------------------------------------------------------------ stages 1..L-1
                for (int s = s_next; s < L; ++s) {
                    const float* W = acquire_blob(t, s);
                    float pre_a = 0.f, pre_b = 0.f;
                    if (it_y >= 0) { pre_a = pre[pa_idx + s * pl.RA4 * BT]; pre_b = pre[pa_idx + s * pl.RA4 * BT + BT]; }
                    WN_TICK(4);
                    {
                        const int e0 = pl.ex_yx + (s - 1) * YX;
                        const uint32_t tag = tagbase + wn_eid_yx(s - 1);
                        if (pp.gate_cycles > 0) { while (clock64() - t_pub < pp.gate_cycles) {} }
                        if (s >= 2) poll_vec2<EG, ER>(xin, e0, G2, yr, e0 + G2, R, xr, tag);
                        else poll_vec<EG>(xin, e0, G2, tag, yr);      // x_0 is already in registers
                    }
                    WN_TICK(0);
                    float* xst = xs + (size_t)(s & 1) * R * BT;
                    float* r1 = red1 + (size_t)(s & 1) * pl.red1_floats;
                    stash<ER>(xst, R, xr);
                    stash<EG>(ys + (size_t)(s & 1) * G2 * BT, G2, yr);
                    // gate pre-activations of layer s (quads [0,NQ_A)) and residual rows x_s (quads [NQ_A,nqc))
                    for (int q = 0; q < nqc; q += NA) {
                        float acc[NA][NV];
#pragma unroll
                        for (int h = 0; h < NA; ++h) {
#pragma unroll
                            for (int v = 0; v < NV; ++v) acc[h][v] = 0.f;
                            const int qq = q + h;
                            if (qq < NQ_A) {
                                quad_fma<EG>(W + pl.lb_Zy + (size_t)qq * G2 * 4, G2, yr, acc[h]);
                                quad_fma<ER>(W + pl.lb_Zx + (size_t)qq * R * 4, R, xr, acc[h]);
                            } else if (qq < nqc) {
                                quad_fma<EG>(W + pl.lb_Xo + (size_t)(qq - NQ_A) * G2 * 4, G2, yr, acc[h]);
                            }
                        }
                        reduce_scatter_multi<NA, NV>(acc, lane);
#pragma unroll
                        for (int h = 0; h < NA; ++h)
                            if (q + h < nqc) quad_store(acc[h], q + h, r1);
                    }
                    WN_TICK(1);
                    if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                    if (gt == 0) {          // stash complete (the barrier ordered every thread's writes)
                        __threadfence_block();
                        *s_stash_cnt = nstash + 1;
                    }
                    ++nstash;
                    WN_TICK(2);
                    const uint32_t tag = tagbase + wn_eid_yx(s);
                    if (it_y >= 0) {
                        const int fr = it_y / BT, fb = it_y % BT;
                        const float a = red_sum(r1, 2 * fr, fb) + pre_a;
                        const float g = red_sum(r1, 2 * fr + 1, fb) + pre_b;
                        publish(pl.ex_yx + s * YX + y0 + fr, fb, cp_y, gate(a, g), tag);
                    }
                    if (it_x >= 0) {
                        // modules.py:160-162  x_s = (conv1x1_out(y_{s-1}) + x_{s-1}) * sqrt(0.5)
                        const int fr = it_x / BT, fb = it_x % BT;
                        const float o = red_sum(r1, NQ_A * 4 + fr, fb) + W[pl.lb_xb + fr];
                        publish(pl.ex_yx + s * YX + G2 + x0r + fr, fb, cp_x, (o + xst[(x0r + fr) * BT + fb]) * RSQRT2, tag);
                    }
                    release_blob(t, s);
                    t_pub = clock64();
                    // the stash of stage s+1 reuses the buffer of stage s-1: the deferred group must be done with it
                    if (s >= 2) { wait_count(s_ddone_cnt, WN_GW * (ndone + 1), 0x08000000u); ++ndone; }
                    WN_TICK(3);
                }
                if (step_dead) break;
                bool tail_done = false;
                if constexpr (LEAN_T) {
                    if (lean) {
                        // ---- lean stage L, head 1, head 2 (same arithmetic as the generic code below)
                        tail_done = true;
                        const float* H = acquire_lean(t, L);
                        lean_poll_x(l_in0 + (long long)(L - 1) * l_inc, tagbase + (uint32_t)(L - 1));
                        lean_poll_y(l_in0 + (long long)(L - 1) * l_inc, tagbase + (uint32_t)(L - 1));
                        WN_TICK(0);
                        float* xst = xs + (L & 1) * R;
#pragma unroll
                        for (int j = 0; j < LX; ++j)
                            *reinterpret_cast<float2*>(xst + 2 * gt + 2 * WN_NTC * j) = make_float2(xr[2 * j][0], xr[2 * j + 1][0]);
                        float* r1 = red1 + (L & 1) * pl.red1_floats;
                        float* r1a = red1 + ((L + 1) & 1) * pl.red1_floats;
                        {
                            float a4[4] = {0.f, 0.f, 0.f, 0.f};
                            lean_quad_y(H + pl.tb_Sk + 8 * gt, a4);
                            reduce_scatter<4>(a4, lane);
                            if ((lane & 7) == 0) r1[(lane >> 3) * WN_GW + gw] = a4[0];
                        }
                        WN_TICK(1);
                        if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                        if (gt == 0) {
                            fence_cta();
                            *s_stash_cnt = nstash + 1;
                        }
                        ++nstash;
                        // skip rows of layers 0..L-2 were accumulated by the deferred group: wait for its stage L-1
                        count_lean(s_ddone_cnt, WN_GW * (ndone + 1), 0u, 0x08000001u);
                        ++ndone;
                        WN_TICK(2);
                        if (it_s >= 0) {
                            float tot = quad_sum(r1 + it_s * WN_GW) + H[pl.tb_sb + it_s];
                            tot = skipacc[it_s] + tot;
                            st_pair(l_pubs, fmaxf(tot * pl.skip_scale, 0.f), tagbase + (uint32_t)L);
                        }
                        WN_TICK(3);
                        if (na > 0) {
                            lean_poll_y(l_sk0, tagbase + (uint32_t)L);
                            float a4[4] = {0.f, 0.f, 0.f, 0.f};
                            lean_quad_y(H + pl.tb_Ha + 8 * gt, a4);
                            reduce_scatter<4>(a4, lane);
                            if ((lane & 7) == 0) r1a[(lane >> 3) * WN_GW + gw] = a4[0];
                        }
                        if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                        if (it_a >= 0)
                            st_pair(l_puba, fmaxf(quad_sum(r1a + it_a * WN_GW) + H[pl.tb_Hab + it_a], 0.f), tagbase + (uint32_t)L + 1u);
                        if (nb > 0) {
                            lean_poll_y(l_h10, tagbase + (uint32_t)L + 1u);
                            float a4[4] = {0.f, 0.f, 0.f, 0.f};
                            lean_quad_y(H + pl.tb_Hb + 8 * gt, a4);
                            reduce_scatter<4>(a4, lane);
                            if ((lane & 7) == 0) r1[(lane >> 3) * WN_GW + gw] = a4[0];
                        }
                        if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                        if (it_b >= 0)
                            st_pair(l_pubb, quad_sum(r1 + it_b * WN_GW) + H[pl.tb_Hbb + it_b], tagbase + (uint32_t)L + 2u);
                        release_blob(t, L);
                    }
                }
                if (!tail_done) {
                // ------------------------------------------------------------ stage L: skip of the last layer
                const float* H = acquire_blob(t, L);
                {
                    const int e0 = pl.ex_yx + (L - 1) * YX;
                    const uint32_t tag = tagbase + wn_eid_yx(L - 1);
                    if (L >= 2) poll_vec2<EG, ER>(xin, e0, G2, yr, e0 + G2, R, xr, tag);
                    else poll_vec<EG>(xin, e0, G2, tag, yr);
                }
                WN_TICK(0);
                stash<ER>(xs + (size_t)(L & 1) * R * BT, R, xr);
                float* r1 = red1 + (size_t)(L & 1) * pl.red1_floats;
                gemv<EG>(H + pl.tb_Sk, pl.NQ_BS, G2, yr, r1);
                WN_TICK(1);
                if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                if (gt == 0) {
                    __threadfence_block();
                    *s_stash_cnt = nstash + 1;
                }
                ++nstash;
                // skip rows of layers 0..L-2 were accumulated by the deferred group: wait for its stage L-1
                if (L >= 2) { wait_count(s_ddone_cnt, WN_GW * (ndone + 1), 0x08000001u); ++ndone; }
                WN_TICK(2);
                if (it_s >= 0) {
                    // (s_0 + ... + s_{L-2}) + s_{L-1}, * sqrt(1/L), first ReLU of the head (wavenet.py:312-315)
                    const int fr = it_s / BT, fb = it_s % BT;
                    float tot = red_sum(r1, fr, fb) + H[pl.tb_sb + fr];
                    if (L >= 2) tot = skipacc[it_s] + tot;
                    publish(pl.ex_sk + s0 + fr, fb, cp_s, fmaxf(tot * pl.skip_scale, 0.f), tagbase + wn_eid_sk(pl));
                }
                WN_TICK(3);
                // ---------------------------------------------------------------- head (wavenet.py:315-319)
                float* r1a = red1 + (size_t)((L + 1) & 1) * pl.red1_floats;
                if (na > 0) {
                    WN_DISPATCH_E(ES, { float h[E][BT];
                                        poll_vec<E>(xin, pl.ex_sk, S, tagbase + wn_eid_sk(pl), h);
                                        gemv<E>(H + pl.tb_Ha, pl.NQ_HA, S, h, r1a); });
                }
                if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                if (it_a >= 0) {
                    const int fr = it_a / BT, fb = it_a % BT;
                    publish(pl.ex_h1 + a0 + fr, fb, cp_a, fmaxf(red_sum(r1a, fr, fb) + H[pl.tb_Hab + fr], 0.f), tagbase + wn_eid_h1(pl));
                }
                if (nb > 0) {
                    WN_DISPATCH_E(ES, { float h[E][BT];
                                        poll_vec<E>(xin, pl.ex_h1, S, tagbase + wn_eid_h1(pl), h);
                                        gemv<E>(H + pl.tb_Hb, pl.NQ_HB, S, h, r1); });
                }
                if (bar_or_n<1, WN_NTC>(dead)) { step_dead = true; break; }
                if (it_b >= 0) {
                    const int fr = it_b / BT, fb = it_b % BT;
                    publish(pl.ex_h2 + b0 + fr, fb, cp_b, red_sum(r1, fr, fb) + H[pl.tb_Hbb + fr], tagbase + wn_eid_h2(pl));
                }
                release_blob(t, L);
                }
                {
                    WN_DISPATCH_E(EO, { float h[E][BT];
                                        poll_vec<E>(xin, pl.ex_h2, O, tagbase + wn_eid_h2(pl), h);
                                        stash<E>(hs, O, h); });
                }
                WN_TICK(5);
            } while (false);
            // ---- both groups meet: sampler (one warp per utterance), then the next step
            if (bar_groups(dead || step_dead)) return;
            // the deferred group finished its stage L before this barrier
            if (L >= 1) ++ndone;
            step_tail(t);
            if (bar_groups(false)) return;
            WN_TICK(6);
        }
        if (prof) {
            for (int i = 0; i < 8; ++i) pp.prof[(size_t)p * 16 + i] = pc[i];
        }
#undef WN_TICK
    }

    // head outputs are in hs: optional dump, sampling, feedback for the next step (all 8 compute warps)
    __device__ __forceinline__ void step_tail(int t) {
        const int O = pl.O, T = pp.T;
        if (p == 0 && pp.params_out != nullptr) {
            for (int i = tid; i < O * BT; i += WN_NT) {
                const int o = i / BT, b = i % BT;
                if (b < pp.B) pp.params_out[((size_t)b * O + o) * T + t] = hs[i];
            }
            // the softmax sampler overwrites hs in place: finish the copy first (block-uniform)
            if (pl.head_kind == 2) bar_groups(false);
        }
        if (warp < BT) {
            sample_utt(t, warp);
#ifndef WN_NO_SAMPLER_SYNCWARP
            __syncwarp();                 // every lane has read step t's draws before they are overwritten
#endif
            if (t + 1 < T) fetch_noise(t + 1, warp);
        }
        // advance the ring positions to (t+1) mod delay
        for (int i = WN_NT - 1 - tid; i < pl.L * (pl.kw - 1); i += WN_NT) {
            const int pos = ringtab[i * 3 + 2] + 1;
            ringtab[i * 3 + 2] = (pos == ringtab[i * 3 + 1]) ? 0 : pos;
        }
    }

    // --------------------------------------------------------------------------------------
    // deferred group (warps 4-7): queued older-tap products and skip rows, one stage behind
    // --------------------------------------------------------------------------------------
    __device__ void def_loop() {
        const int L = pl.L, R = pl.R, G2 = pl.G2, T = pp.T, P = pl.P, kw = pl.kw;
        int s0, ns;
        wn_part(pl.S, P, p, s0, ns);
        const int NQ_D = pl.NQ_D, NQ_BS = pl.NQ_BS, nqd = (kw > 1 ? NQ_D : 0) + NQ_BS;
        const int qoff = (kw > 1 ? NQ_D : 0);
        const int nring_items = (kw - 1) * pl.RA * BT, nskip_items = ns * BT;
        float xr[ER][BT], yr[EG][BT];
        const bool prof = (pp.prof != nullptr) && gt == 0;
        long long pc[4] = {0, 0, 0, 0}, tc = 0;
#define WN_TICK(i) if (prof) { const long long now_ = clock64(); pc[i] += now_ - tc; tc = now_; }
        int nstash = 0;

        // lean variant of a deferred stage (see crit_loop): two older-tap quads (acc[0..7]) + one skip quad (accs)
        constexpr bool LEAN_T = LEAN;
        constexpr bool lean = LEAN;
        const int d_tap = (gt < nring_items) ? gt / pl.RA : 0, d_rr = (gt < nring_items) ? gt % pl.RA : 0;
        // the residual stream is published from here (threads 64.. of the group), half a stage before the critical
        // group of any block asks for it: its poll then spins on y alone (a third of the exchange)
        int d_x0r, d_nx;
        wn_part(R, P, p, d_x0r, d_nx);
        const int d_fx = (gt >= 64 && gt < 64 + d_nx) ? gt - 64 : -1;
        uint2* d_pubx = pp.xbuf;
        long long d_inc = 0;
        if (lean) {
            d_pubx = pp.xbuf + wn_pair_index((long long)pl.ex_yx) + wn_pair_index((long long)(G2 + d_x0r + max(d_fx, 0)));
            d_inc = wn_pair_index((long long)(G2 + R));
        }
        const float RSQRT2 = 0.70710678118654752440f;         // math.sqrt(0.5), modules.py:162
        auto lean_stage = [&](int t, int s, int layer, const float* Td, const float* Sk, const float* skb) -> bool {
            count_lean(s_stash_cnt, nstash + 1, 32u, 0x04000000u);
            ++nstash;
            WN_TICK(0);
            const float* xst = xs + (s & 1) * R;
            const float* yst = ys + (s & 1) * G2;
            float* red = red2 + (s & 1) * pl.red2_floats;
            // ---- phase 1: the block's rows of x_s, published at once (the critical groups of all blocks need them
            // for stage s+1, before anything else this group produces)
            // rows of x_s = (conv1x1_out_{s-1} y_{s-1} + b + x_{s-1}) sqrt(.5)  (modules.py:160-162), not in the tail stage
            float accx[4] = {0.f, 0.f, 0.f, 0.f};
            const float* Wb = Sk ? Sk - pl.lb_Sk : nullptr;     // the blob of stage s
            if (Sk) {
                const float* wx = Wb + pl.lb_Xo + 8 * gt;
#pragma unroll
                for (int j = 0; j < EG / 2; ++j) {
                    const float2 yv2 = *reinterpret_cast<const float2*>(yst + 2 * gt + 2 * WN_NTC * j);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = *reinterpret_cast<const float4*>(wx + (2 * WN_NTC * j + h) * 4);
                        const float yv = h ? yv2.y : yv2.x;
                        accx[0] = fmaf(a.x, yv, accx[0]); accx[1] = fmaf(a.y, yv, accx[1]);
                        accx[2] = fmaf(a.z, yv, accx[2]); accx[3] = fmaf(a.w, yv, accx[3]);
                    }
                }
            }
            if (Sk) {
                reduce_scatter<4>(accx, lane);
                if ((lane & 7) == 0) red[(12 + (lane >> 3)) * WN_GW + gw] = accx[0];
                if (bar_or_n<2, WN_NTC>(dead)) dead = true;
                if (!dead && d_fx >= 0) {
                    const float4 q = *reinterpret_cast<const float4*>(red + (12 + d_fx) * WN_GW);
                    const float o = ((q.x + q.y) + (q.z + q.w)) + Wb[pl.lb_xb + d_fx];
                    st_pair(d_pubx + (long long)s * d_inc, (o + xst[d_x0r + d_fx]) * RSQRT2,
                            (uint32_t)t * ((uint32_t)L + 3u) + 1u + (uint32_t)s);
                }
            }
            // ---- phase 2: queued older-tap products and skip rows
            float acc[8], accs[4];
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) accs[v] = 0.f;
            {
                const float *w0 = Td + 8 * gt, *w1 = Td + (size_t)R * 4 + 8 * gt;
#pragma unroll
                for (int j = 0; j < ER / 2; ++j) {
                    const float2 xv2 = *reinterpret_cast<const float2*>(xst + 2 * gt + 2 * WN_NTC * j);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ko = (2 * WN_NTC * j + h) * 4;
                        const float4 a = *reinterpret_cast<const float4*>(w0 + ko);
                        const float4 b = *reinterpret_cast<const float4*>(w1 + ko);
                        const float xv = h ? xv2.y : xv2.x;
                        acc[0] = fmaf(a.x, xv, acc[0]); acc[1] = fmaf(a.y, xv, acc[1]);
                        acc[2] = fmaf(a.z, xv, acc[2]); acc[3] = fmaf(a.w, xv, acc[3]);
                        acc[4] = fmaf(b.x, xv, acc[4]); acc[5] = fmaf(b.y, xv, acc[5]);
                        acc[6] = fmaf(b.z, xv, acc[6]); acc[7] = fmaf(b.w, xv, acc[7]);
                    }
                }
            }
            if (Sk) {
                const float* ws = Sk + 8 * gt;
#pragma unroll
                for (int j = 0; j < EG / 2; ++j) {
                    const float2 yv2 = *reinterpret_cast<const float2*>(yst + 2 * gt + 2 * WN_NTC * j);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = *reinterpret_cast<const float4*>(ws + (2 * WN_NTC * j + h) * 4);
                        const float yv = h ? yv2.y : yv2.x;
                        accs[0] = fmaf(a.x, yv, accs[0]); accs[1] = fmaf(a.y, yv, accs[1]);
                        accs[2] = fmaf(a.z, yv, accs[2]); accs[3] = fmaf(a.w, yv, accs[3]);
                    }
                }
            }
            reduce_scatter<8>(acc, lane);
            reduce_scatter<4>(accs, lane);
            if ((lane & 3) == 0) red[(lane >> 2) * WN_GW + gw] = acc[0];
            if (Sk != nullptr && (lane & 7) == 0) red[(8 + (lane >> 3)) * WN_GW + gw] = accs[0];
            WN_TICK(1);
            const bool d = bar_or_n<2, WN_NTC>(dead);
            if (!d) {

                if (gt < nring_items) {
                    const int e = (layer * (kw - 1) + d_tap) * 3;
                    const float4 q = *reinterpret_cast<const float4*>(red + gt * WN_GW);
                    ring[((size_t)ringtab[e] + ringtab[e + 2]) * pl.RA4 + d_rr] = (q.x + q.y) + (q.z + q.w);
                }
                if (Sk != nullptr && gt >= 32 && gt < 32 + nskip_items) {
                    const int f = gt - 32;
                    const float4 q = *reinterpret_cast<const float4*>(red + (8 + f) * WN_GW);
                    const float h = ((q.x + q.y) + (q.z + q.w)) + skb[f];
                    skipacc[f] = (layer == 0) ? h : skipacc[f] + h;
                }
            }
            fence_cta();
            __syncwarp();
            if (lane == 0) atomicAdd((int*)s_ddone_cnt, 1);
            WN_TICK(2);
            return d;
        };

        // one deferred stage: `Td` = older taps of `layer` (uses x), `Sk`/`skb` = skip rows of `layer`
        // (uses y; nullptr in the tail stage, where the critical group evaluates them itself)
        auto stage = [&](int t, int s, int layer, const float* Td, const float* Sk, const float* skb) {
            if constexpr (LEAN_T) {
                if (lean) return lean_stage(t, s, layer, Td, Sk, skb);
            }
            wait_count<true>(s_stash_cnt, nstash + 1, 0x04000000u);
            ++nstash;
            WN_TICK(0);
            unstash<ER>(xs + (size_t)(s & 1) * R * BT, R, xr);
            if (Sk) unstash<EG>(ys + (size_t)(s & 1) * G2 * BT, G2, yr);
            float* red = red2 + (size_t)(s & 1) * pl.red2_floats;
            const int nq = Sk ? nqd : qoff;
            for (int q = 0; q < nq; q += NA) {
                float acc[NA][NV];
#pragma unroll
                for (int h = 0; h < NA; ++h) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[h][v] = 0.f;
                    const int qq = q + h;
                    if (qq < qoff) quad_fma<ER>(Td + (size_t)qq * R * 4, R, xr, acc[h]);
                    else if (qq < nq) quad_fma<EG>(Sk + (size_t)(qq - qoff) * G2 * 4, G2, yr, acc[h]);
                }
                reduce_scatter_multi<NA, NV>(acc, lane);
#pragma unroll
                for (int h = 0; h < NA; ++h)
                    if (q + h < nq) quad_store(acc[h], q + h, red);
            }
            WN_TICK(1);
            const bool d = bar_or_n<2, WN_NTC>(dead);
            if (!d) {
                // older-tap products of `layer` -> history ring (consumed at steps t+d, t+2d, conv.py:32-44)
                for (int f = gt; f < nring_items; f += WN_NTC) {
                    const int dr = f / BT, db = f % BT, tap = dr / pl.RA, rr = dr % pl.RA;
                    const int e = (layer * (kw - 1) + tap) * 3;
                    ring[((size_t)ringtab[e] + ringtab[e + 2]) * pl.RA4 * BT + rr * BT + db] = red_sum(red, dr, db);
                }
                // skip rows, accumulated in layer order (wavenet.py:312)
                if (Sk) {
                    for (int f = gt; f < nskip_items; f += WN_NTC) {
                        const int fr = f / BT, fb = f % BT;
                        const float h = red_sum(red, qoff * 4 + fr, fb) + skb[fr];
                        skipacc[f] = (layer == 0) ? h : skipacc[f] + h;
                    }
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd((int*)s_ddone_cnt, 1);
            WN_TICK(2);
            return d;
        };

        build_pre(0);
        if (bar_groups(dead)) return;
        for (int t = 0; t < T; ++t) {
            if (prof) tc = clock64();
            bool step_dead = false;
            release_blob(t, 0);                       // stage 0 has no deferred work
            for (int s = 1; s < L && !step_dead; ++s) {
                const float* W = acquire_blob(t, s);
                step_dead = stage(t, s, s - 1, W + pl.lb_Td, W + pl.lb_Sk, W + pl.lb_sb);
                release_blob(t, s);
            }
            if (!step_dead) {
                const float* H = acquire_blob(t, L);
                step_dead = stage(t, L, L - 1, H + pl.tb_Td, nullptr, nullptr);
                release_blob(t, L);
            }
            if (!step_dead && t + 1 < T) {
                // all rings of step t are written (group barrier inside stage()): pre-sums of step t+1
                if (bar_or_n<2, WN_NTC>(dead)) step_dead = true;
                else build_pre(t + 1);
            }
            if (bar_groups(dead || step_dead)) return;
            step_tail(t);
            if (bar_groups(false)) return;
            WN_TICK(3);
        }
        if (prof) {
            for (int i = 0; i < 4; ++i) pp.prof[(size_t)p * 16 + 8 + i] = pc[i];
        }
#undef WN_TICK
    }
};

// ------------------------------------------------------------------------------------------
// kernel entry
// ------------------------------------------------------------------------------------------
template <int BT, int ER, int EG, bool LEAN = false>
__global__ void __launch_bounds__(WN_NTHREADS, 1)
wn_persistent_kernel(const __grid_constant__ WnPlan pl, const __grid_constant__ WnPtrs pp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Engine<BT, ER, EG, LEAN> eng(pl, pp, smem_raw);
    const int tid = eng.tid, p = blockIdx.x;      // logical thread index (see Engine)
    const int nslots = pl.nres + pl.nring;
    if (tid == 0) {
        for (int i = 0; i < nslots; ++i) mbar_init(&eng.bar_full[i], 1);
        for (int i = 0; i < pl.nring; ++i) mbar_init(&eng.bar_empty[i], WN_NWARP);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&eng.bar_cfull[i], 1);
            mbar_init(&eng.bar_cempty[i], WN_GW);
        }
        *eng.s_abort = 0;
        *eng.s_stash_cnt = 0;
        *eng.s_ddone_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the history (== the reference's zero-initialised queue, conv.py:35-36) and scratch
    if (pl.ring_in_smem) {
        const size_t n = (size_t)pl.ring_pos_total * pl.RA4 * BT;
        for (size_t i = tid; i < n; i += WN_NTHREADS) eng.ring[i] = 0.f;
    }
    for (int i = tid; i < pl.NSm * BT + 4; i += WN_NTHREADS) eng.skipacc[i] = 0.f;
    for (int i = tid; i < pl.L * (pl.kw - 1); i += WN_NTHREADS) {
        eng.ringtab[i * 3] = pp.ringtab[i * 2];           // offset of the ring (in positions)
        eng.ringtab[i * 3 + 1] = pp.ringtab[i * 2 + 1];   // delay D
        eng.ringtab[i * 3 + 2] = 0;                       // t mod D
    }
    for (int k = tid; k < pl.R; k += WN_NTHREADS) {
        eng.first[k] = (pl.input_kind == 0) ? pp.first_w[k] : 0.f;
        eng.first[pl.R + k] = pp.first_b[k];
    }
    {
        // static part of the pre-activation: (folded) conv bias + global-conditioning projection
        // (modules.py:148-152 recomputes Wg.g every step although g is constant; fold it once)
        int y0, ny;
        wn_part(pl.G2, pl.P, p, y0, ny);
        const float* blob0 = pp.wpack + (size_t)p * pl.cta_w_floats;
        const int n = pl.L * pl.RA4 * BT;
        for (int i = tid; i < n; i += WN_NTHREADS) {
            const int b = i % BT, rr = (i / BT) % pl.RA4, l = i / (BT * pl.RA4);
            float v = 0.f;
            if (rr < pl.RA && (rr >> 1) < ny) {
                v = blob0[wn_blob_off(pl, l) + (l == 0 ? pl.fb_zb : pl.lb_zb) + rr];
                if (pp.gbias != nullptr && b < pp.B) {
                    const int grow = (rr & 1) ? pl.G2 + y0 + (rr >> 1) : y0 + (rr >> 1);
                    v += pp.gbias[((size_t)b * pl.L + l) * pl.G + grow];
                }
            }
            eng.sb[i] = v;
        }
    }
    // feedback for step 0 (wavenet.py:281-301)
    if (tid < BT) {
        const int b = tid;
        float v = 0.f;
        int idx = -1;
        if (b < pp.B) {
            if (pl.input_kind == 0) {
                if (pp.T_test > 0) v = pp.test_scalar[(size_t)b * pp.T_test];
                else if (pp.initial) v = pp.initial[b];
            } else {
                if (pp.T_test > 0) idx = pp.test_index ? pp.test_index[(size_t)b * pp.T_test] : -1;
                else if (pp.initial_dense) idx = -1;
                else if (pp.initial_rows) idx = pp.initial_rows[b];
                else idx = pp.initial_index;
            }
        } else if (pl.input_kind != 0) idx = 0;
        eng.s_in[b] = v;
        eng.s_idx[b] = idx;
    }
    if (pl.input_kind != 0) {
        const float* dsrc = nullptr;
        size_t stride = 0;
        if (pp.T_test > 0 && pp.test_dense != nullptr) { dsrc = pp.test_dense; stride = (size_t)pp.T_test * pl.O; }
        else if (pp.T_test == 0 && pp.initial_dense != nullptr) { dsrc = pp.initial_dense; stride = (size_t)pl.O; }
        for (int i = tid; i < BT * pl.O; i += WN_NTHREADS) {
            const int b = i / pl.O, o = i % pl.O;
            eng.s_dense[i] = (dsrc && b < pp.B) ? dsrc[(size_t)b * stride + o] : 0.f;
        }
    }
    const int warp = tid >> 5;
    if (warp < BT && warp < WN_NWARP) eng.fetch_noise(0, warp);
    __syncthreads();
    if (warp == WN_NWARP) {
        eng.tma_loop();
        return;
    }
    if (warp == WN_NWARP + 1) {
        if (pl.C > 0) eng.cond_loop();
        return;
    }
    if (warp < WN_GW) eng.crit_loop();
    else eng.def_loop();
}

// gbias[b][l][row] = Wg_l[row,:] . g_b   (modules.py:148-152), once per call
__global__ void wn_gbias_kernel(const float* __restrict__ wg, const float* __restrict__ g, float* __restrict__ out,
                                int L, int G, int gin) {
    const int l = blockIdx.x, b = blockIdx.y;
    for (int row = threadIdx.x; row < G; row += blockDim.x) {
        const float* w = wg + ((size_t)l * G + row) * gin;
        float a = 0.f;
        for (int i = 0; i < gin; ++i) a = fmaf(w[i], g[(size_t)b * gin + i], a);
        out[((size_t)b * L + l) * G + row] = a;
    }
}

// stand-alone samplers over (B,O,T): the reference's mixture.py entry points
__global__ void wn_sample_kernel(const float* __restrict__ y, int B, int O, int T, const float* __restrict__ u1,
                                 const float* __restrict__ n2, float* __restrict__ out, int gauss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int b = i / T, t = i % T;
    const float* yb = y + (size_t)b * O * T + t;
    float mean, ls;
    const int K = (O == 2) ? 1 : O / 3;
    if (K > 1 || (!gauss)) {
        float best = -INFINITY;
        int bi = 0;
        for (int k = 0; k < K; ++k) {
            const float gk = yb[(size_t)k * T] - logf(-logf(u1[((size_t)t * B + b) * K + k]));
            if (gk > best) {
                best = gk;
                bi = k;
            }
        }
        mean = yb[(size_t)(K + bi) * T];
        ls = yb[(size_t)(2 * K + bi) * T];
    } else if (O == 2) {
        mean = yb[0];
        ls = yb[(size_t)T];
    } else {
        mean = yb[(size_t)T];
        ls = yb[(size_t)2 * T];
    }
    const float v = n2[(size_t)t * B + b];
    float xv;
    if (!gauss) xv = __fadd_rn(mean, __fmul_rn(expf(ls), __fsub_rn(logf(v), logf(__fsub_rn(1.0f, v)))));
    else xv = __fadd_rn(__fmul_rn(v, expf(ls)), mean);
    out[i] = fminf(fmaxf(xv, -1.0f), 1.0f);
}

}  // namespace wn
