// wn6_kernel.cuh — the persistent sm_100a synthesis kernel (thread-block clusters + DSMEM).
//
// One launch == one WaveNet.incremental_forward() call (reference wavenet.py:215-343): the whole T-step
// loop, including the sampler, runs on the device.  The grid is NC clusters of CS blocks (wn6_plan.h);
// every stage of a step is decomposed over (cluster = output rows) x (rank = K-slice of the input).
//
// Stages of one generated sample (the algebra is the reference's, re-associated on the host):
//   stage 0      : x_0 (first 1x1 conv of the fed-back sample, wavenet.py:308; every block evaluates its
//                  K-slice locally) -> current tap of layer 0 -> tanh*sigmoid -> publish y_0
//   stage s < L  : from (y_{s-1}, x_{s-1}):  z_s = M_{s-1} y_{s-1} + V_s x_{s-1} + bias + conditioning + queued
//                  older taps, with V_s = sqrt(.5) W_s[:,:,kw-1] and M_{s-1} = V_s Wo_{s-1} folded on the host
//                  (conv1x1_out of layer s-1 rides inside the current tap of layer s: ONE exchange per layer),
//                  x_s = (Wo_{s-1} y_{s-1} + bo + x_{s-1}) sqrt(.5)  (modules.py:160-162) -> publish (y_s, x_s).
//                  Deferred (off the critical path): the OLDER taps' products W_{s-1}[:,:,k<kw-1] x_{s-1}(t), queued
//                  for steps t+d, t+2d (replaces the input shift register of conv.py:32-44 by a queue of output
//                  partials), and conv1x1_skip_{s-1}, accumulated in layer order (wavenet.py:312).
//   stage L      : skip rows of the last layer -> total skip * sqrt(1/L) -> ReLU -> publish
//   stage L+1,+2 : last_conv_layers (wavenet.py:315-319)
// then every block reads the O head outputs and evaluates the sampler (mixture.py) redundantly from identical
// noise, so the sample itself needs no broadcast.
//
// Warp roles (15 warps):
//   pollers  (2) : poll the block's K-slice of the previous stage's tagged pairs in L2 (ld.relaxed.gpu, 16 bytes
//                  = 2 pairs per load) into the stage-input buffer in shared memory; run the sampler.
//   compute  (8) : passes (wn6_plan.h): weight tile from shared memory x stage input -> 16-lane butterfly ->
//                  partial sums to the row owners through DSMEM (st.async ... mbarrier::complete_tx::bytes).
//   F0, F1       : owner-side finalisers: wait for the CS partials (mbarrier tx count), add bias / pre-sums / residual,
//                  gate, publish the block's values with st.relaxed.gpu (value and tag in one 8-byte word: no fence).
//   DF           : deferred finaliser: history rings, skip accumulator, next step's pre-sum table.
//   TMA          : streams the packed weight blobs global -> shared with cp.async.bulk + mbarrier (SASS UBLKCP).
//   COND         : local-conditioning projection of the owner's gate rows, one step ahead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "wn6_plan.h"

struct Wn6Ptrs {
    const float* wpack;        // [P][cta_w_floats]
    const float* cwpack;       // [P][cta_cw_floats]
    const float* bpack;        // [P][cta_b_floats]
    const Wn6Pass* passes;     // [npass]
    const float* gbias;        // [B][L][G] = Wg_l . g_b   (NULL without global conditioning)
    const float* first_w;      // scalar input: [R];  one-hot input: transposed [O][R]
    const float* first_b;      // [R]
    uint2* xbuf;               // exchange pairs
    float* ring_g;             // [P][ring floats] when the history rings do not fit in shared memory
    const int* ringtab;        // [L*(kw-1)*2] : (offset in positions, delay D)
    int* err;                  // [4] device fault word, what, block, thread
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
    long long* prof;           // optional [P][16] cycle counters
};

#define WN6_FLAG_SOFTMAX 1u
#define WN6_FLAG_QUANTIZE 2u

namespace wn6 {

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 ld_pair2(const uint2* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pair(uint2* p, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ int ld_flag(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// shared::cluster address of `addr` (a shared::cta address of this block) in block `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// remote shared-memory store whose completion is counted (in bytes) on a remote mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
                 "r"(__float_as_uint(v)), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_f32x2(uint32_t raddr, float v0, float v1, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1,%2}, [%3];" ::"r"(raddr),
                 "r"(__float_as_uint(v0)), "r"(__float_as_uint(v1)), "r"(rbar)
                 : "memory");
}
// arrive on an mbarrier of another block of the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// named barrier `ID` over `N` threads
template <int ID, int N>
__device__ __forceinline__ void bar_sync_n() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// Philox4x32-10 (counter-based; the same (seed, step, utterance, slot) gives the same draw in every block, which
// is what lets all blocks sample redundantly)
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
__device__ __noinline__ bool check_abort_slow(volatile int* s_abort, int* err, long long timeout, uint32_t what, int p,
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

// butterfly over the 16 lanes of a row quad: every level that still has more than one value also halves
// the value set.  Afterwards lane `sub` holds value(s) [sub*NV/16, ...) (NV >= 16) or value sub*NV/16 (NV < 16).
template <int NV>
__device__ __forceinline__ void reduce16(float (&v)[NV], int lane) {
    int n = NV;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
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

template <int BT>
struct Engine {
    static constexpr int NV = 4 * BT;
    const Wn6Plan& pl;
    const Wn6Ptrs& pp;
    unsigned char* sm;
    int tid, warp, lane, p, c, rank;
    uint64_t *bar_full, *bar_empty, *bar_cfull, *bar_cempty, *bar_in, *bar_free, *bar_part, *bar_dpart, *bar_partx,
        *bar_dfree, *bar_pre, *bar_x0, *bar_ps;
    volatile int* s_abort;
    volatile int* s_ddone;         // deferred stages completed by DF (monotonic)
    Wn6Pass* passes;
    int* ringtab;                  // [e][3]: offset, delay, t mod delay
    float *xin, *part, *dpart, *partx, *sb, *pre, *cond, *bias, *skipacc, *xown, *x0own, *hs, *noise, *x0w, *slots;
    int* h2map;
    volatile float* ring;
    float* s_in;      // [BT] scalar feedback
    int* s_idx;       // [BT] class feedback
    float* s_dense;   // [BT][O] dense feedback
    bool dead;

    __device__ Engine(const Wn6Plan& pl_, const Wn6Ptrs& pp_, unsigned char* sm_) : pl(pl_), pp(pp_), sm(sm_) {
        tid = threadIdx.x;
        warp = tid >> 5;
        lane = tid & 31;
        p = blockIdx.x;
        rank = (int)cluster_ctarank();
        c = p / pl.CS;
        const int nslots = pl.nres + pl.nring;
        bar_full = reinterpret_cast<uint64_t*>(sm + pl.sm_bar);
        bar_empty = bar_full + nslots;
        bar_cfull = bar_empty + (pl.nring > 0 ? pl.nring : 1);
        bar_cempty = bar_cfull + 2;
        bar_in = bar_cempty + 2;
        bar_free = bar_in + 2;
        bar_part = bar_free + 2;
        bar_dpart = bar_part + 2;
        bar_partx = bar_dpart + 2;
        bar_dfree = bar_partx + 2;
        bar_pre = bar_dfree + 2;
        bar_x0 = bar_pre + 1;
        bar_ps = bar_x0 + 1;
        s_abort = reinterpret_cast<volatile int*>(sm + pl.sm_misc);
        s_ddone = s_abort + 1;
        passes = reinterpret_cast<Wn6Pass*>(sm + pl.sm_pass);
        ringtab = reinterpret_cast<int*>(sm + pl.sm_ringtab);
        xin = reinterpret_cast<float*>(sm + pl.sm_xin);
        part = reinterpret_cast<float*>(sm + pl.sm_part);
        dpart = reinterpret_cast<float*>(sm + pl.sm_dpart);
        partx = reinterpret_cast<float*>(sm + pl.sm_partx);
        sb = reinterpret_cast<float*>(sm + pl.sm_sb);
        pre = reinterpret_cast<float*>(sm + pl.sm_pre);
        cond = reinterpret_cast<float*>(sm + pl.sm_cond);
        bias = reinterpret_cast<float*>(sm + pl.sm_bias);
        skipacc = reinterpret_cast<float*>(sm + pl.sm_skipacc);
        xown = reinterpret_cast<float*>(sm + pl.sm_xown);
        x0own = xown + 4 * pl.qB * BT;
        hs = reinterpret_cast<float*>(sm + pl.sm_hs);
        h2map = reinterpret_cast<int*>(hs + pl.O * BT);
        noise = reinterpret_cast<float*>(sm + pl.sm_noise);
        s_in = reinterpret_cast<float*>(sm + pl.sm_in);
        s_idx = reinterpret_cast<int*>(s_in + BT);
        s_dense = reinterpret_cast<float*>(s_idx + BT);
        x0w = reinterpret_cast<float*>(sm + pl.sm_x0w);
        slots = reinterpret_cast<float*>(sm + pl.sm_slots);
        if (pl.ring_in_smem)
            ring = reinterpret_cast<volatile float*>(sm + pl.sm_ring);
        else
            ring = pp.ring_g + (size_t)p * pl.ring_pos_total * 4 * pl.qA * BT;
        dead = false;
    }

    // ---- watchdog: a stuck wait sets the device fault word and makes every block unwind
    __device__ __forceinline__ bool check_abort(uint32_t what, long long& t0) {
        return check_abort_slow(s_abort, pp.err, pp.timeout_cycles, what, p, t0);
    }
    // Waits are WARP-COLLECTIVE (all 32 lanes call them together) and return a warp-uniform verdict, so that a
    // watchdog abort never leaves some lanes of a warp behind in a later shuffle or vote.  `relaxed` waits
    // (anything off the critical path) back off with nanosleep.
    template <bool relaxed = false>
    __device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, uint32_t what) {
        if (!dead) {
            uint32_t spins = 0;
            long long t0 = 0;
            while (!mbar_try_wait(bar, parity)) {
                if (relaxed) __nanosleep(64);
                if (((++spins) & (relaxed ? 63u : 255u)) == 0 && check_abort(what, t0)) {
                    dead = true;
                    break;
                }
            }
        }
        dead = __any_sync(0xffffffffu, dead);
        return !dead;
    }
    // single-lane variant (the TMA lane)
    template <bool relaxed = false>
    __device__ __forceinline__ bool wait_bar_lane(uint64_t* bar, uint32_t parity, uint32_t what) {
        if (dead) return false;
        uint32_t spins = 0;
        long long t0 = 0;
        while (!mbar_try_wait(bar, parity)) {
            if (relaxed) __nanosleep(64);
            if (((++spins) & (relaxed ? 63u : 255u)) == 0 && check_abort(what, t0)) {
                dead = true;
                return false;
            }
        }
        return true;
    }
    __device__ __forceinline__ void wait_count(volatile int* cnt, int need, uint32_t what) {
        if (!dead) {
            uint32_t spins = 0;
            long long t0 = 0;
            while (*cnt < need) {
                if (((++spins) & 255u) == 0 && check_abort(what, t0)) {
                    dead = true;
                    break;
                }
            }
            __threadfence_block();
        }
        dead = __any_sync(0xffffffffu, dead);
    }
    // barrier over the polling warps with a watchdog (a plain bar.sync would hang if one of them aborted)
    uint32_t ps_par = 0;
    __device__ __forceinline__ void poller_sync() {
        if (!dead) mbar_arrive(bar_ps);
        wait_bar(bar_ps, ps_par, 0x00200000u);
        ps_par ^= 1u;
    }
    __device__ __forceinline__ void publish(long long pair, float v, uint32_t tag) { st_pair(pp.xbuf + pair, v, tag); }

    // ======================================================================================
    // weight streaming warp
    // ======================================================================================
    __device__ void tma_loop() {
        if (lane != 0) return;
        const float* base = pp.wpack + (size_t)p * pl.cta_w_floats;
        for (int i = 0; i < pl.nres; ++i) {
            const uint32_t bytes = (uint32_t)wn6_blob_floats(pl, i) * 4u;
            mbar_expect_tx(&bar_full[i], bytes);
            bulk_g2s(slots + (size_t)i * pl.slot_floats, base + wn6_blob_off(pl, i), bytes, &bar_full[i]);
        }
        const int nstream = pl.nblobs - pl.nres;
        if (nstream <= 0) return;
        const uint32_t total = (uint32_t)pp.T * (uint32_t)nstream;
        int i = pl.nres;
        uint32_t s = 0, u = 0;
        for (uint32_t js = 0; js < total; ++js) {
            if (u > 0) {
                if (!wait_bar_lane<true>(&bar_empty[s], (u - 1) & 1u, 0x40000000u | s)) return;
            }
            const uint32_t bytes = (uint32_t)wn6_blob_floats(pl, i) * 4u;
            uint64_t* fb = &bar_full[pl.nres + s];
            mbar_expect_tx(fb, bytes);
            bulk_g2s(slots + (size_t)(pl.nres + s) * pl.slot_floats, base + wn6_blob_off(pl, i), bytes, fb);
            if (++i == pl.nblobs) i = pl.nres;
            if (++s == (uint32_t)pl.nring) { s = 0; ++u; }
        }
    }

    // ======================================================================================
    // conditioning warp: cond[t&1][l][row][b] = Wc_l[own gate rows] . c_t  (modules.py:141-145), one step
    // ahead; the weights come straight from L2 (they are read once per step)
    // ======================================================================================
    __device__ void cond_loop() {
        const int C = pl.C, L = pl.L, T = pp.T, B = pp.B, RA4 = 4 * pl.qA;
        constexpr int M = ilog2c(NV);
        const float* cw = pp.cwpack + (size_t)p * pl.cta_cw_floats;
        for (int t = 0; t < T; ++t) {
            const int par = t & 1, u = t >> 1;
            if (u > 0) {
                if (!wait_bar<true>(&bar_cempty[par], (u - 1) & 1u, 0x20000000u)) return;
            }
            float ct[BT][WN6_MAX_CI];
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int i = 0; i < WN6_MAX_CI; ++i) {
                    const int ch = lane + 32 * i;
                    ct[b][i] = (b < B && ch < C) ? __ldg(pp.c + ((size_t)b * T + t) * C + ch) : 0.f;
                }
            float* dst = cond + (size_t)par * L * RA4 * BT;
            for (int l = 0; l < L; ++l) {
                for (int q = 0; q < pl.qA; ++q) {
                    float acc[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
                    const float* wq = cw + ((size_t)(l * pl.qA + q) * C) * 4;
#pragma unroll
                    for (int i = 0; i < WN6_MAX_CI; ++i) {
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
                    // full-warp butterfly (32 lanes): one more level than reduce16
                    {
                        int n = NV;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            if (n > 1) {
                                n >>= 1;
                                const bool hi = (lane & off) != 0;
#pragma unroll
                                for (int i = 0; i < NV / 2; ++i) {
                                    if (i < n) {
                                        const float send = hi ? acc[i] : acc[i + n];
                                        const float keep = hi ? acc[i + n] : acc[i];
                                        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                                    }
                                }
                            } else {
                                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], off);
                            }
                        }
                    }
                    if ((lane & ((32 >> M) - 1)) == 0) {
                        const int v = lane >> (5 - M);   // = row_in_quad*BT + b
                        dst[((size_t)l * RA4 + q * 4 + v / BT) * BT + (v % BT)] = acc[0];
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
            const bool softmax = (pp.flags & WN6_FLAG_SOFTMAX) != 0, quant = (pp.flags & WN6_FLAG_QUANTIZE) != 0;
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
    // pollers
    // ======================================================================================
    static constexpr int NPL = 32 * WN6_NPW;
    // copy `npairs` tagged pairs starting at `src` into dst[0..npairs) once every tag equals `tag`
    __device__ void poll_pairs(const uint2* __restrict__ src, int npairs, uint32_t tag, float* __restrict__ dst, int pl_) {
        const int nld = (npairs + 1) >> 1;
        for (int j0 = pl_; j0 < nld; j0 += 4 * NPL) {
            uint4 q[4];
            uint32_t spins = 0;
            long long t0 = 0;
            while (true) {
                uint32_t bad = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * NPL;
                    if (j < nld) q[u] = ld_pair2(src + 2 * j);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * NPL;
                    if (j < nld) bad |= (q[u].y ^ tag) | ((2 * j + 1 < npairs) ? (q[u].w ^ tag) : 0u);
                }
                if (bad == 0) break;
                if (((++spins) & 63u) == 0 && check_abort(tag, t0)) {
                    dead = true;
                    break;
                }
            }
            if (dead) break;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * NPL;
                if (j < nld) {
                    dst[2 * j] = __uint_as_float(q[u].x);
                    if (2 * j + 1 < npairs) dst[2 * j + 1] = __uint_as_float(q[u].z);
                }
            }
        }
        dead = __any_sync(0xffffffffu, dead);
    }
    // x_0 = first 1x1 conv of the fed-back sample (wavenet.py:308) at the block's K-slice -> dst[k][b], k < Kx,
    // and at the rows the block owns -> x0own
    __device__ void write_x0(float* __restrict__ dst, int pl_, bool own_too) {
        const int Kx = pl.Kx, O = pl.O, R = pl.R;
        const int n = Kx * BT, nown = 4 * pl.qB * BT;
        for (int i = pl_; i < n + (own_too ? nown : 0); i += NPL) {
            const bool own = i >= n;
            const int k = (own ? i - n : i) / BT, b = (own ? i - n : i) % BT;
            const float* tab = own ? x0w + 2 * Kx : x0w;
            const int kk = own ? 4 * pl.qB : Kx;
            float v;
            if (pl.input_kind == 0) {
                v = fmaf(tab[k], s_in[b], tab[kk + k]);
            } else {
                const int g = __float_as_int(tab[k]);
                v = 0.f;
                if (g >= 0) {
                    const int idx = min(s_idx[b], O - 1);          // class ids are range-checked on the host where it can
                    if (idx >= 0) {
                        v = __ldg(pp.first_w + (size_t)idx * R + g) + tab[kk + k];   // one-hot input: a column gather
                    } else {
                        float a = 0.f;
                        for (int o = 0; o < O; ++o) a = fmaf(__ldg(pp.first_w + (size_t)o * R + g), s_dense[b * O + o], a);
                        v = a + tab[kk + k];
                    }
                }
            }
            if (own) x0own[k * BT + b] = v;
            else dst[k * BT + b] = v;
        }
    }
    // all head outputs of step t -> hs, then the sampler (sets the feedback of step t+1)
    __device__ void read_head_and_sample(int t, int pl_) {
        const int nsl = pl.Kh2 * BT;                       // pairs per rank slice
        const uint32_t tag = (uint32_t)t * (uint32_t)pl.NS + (uint32_t)(pl.L + 2) + 1u;
        const int nld_r = (nsl + 1) >> 1, nld = nld_r * pl.CS;
        for (int j = pl_; j < nld; j += NPL) {
            const int rr = j / nld_r, jj = j % nld_r;
            const uint2* src = pp.xbuf + wn6_ex_off(pl, pl.L + 2, rr) + 2 * jj;
            const bool two = 2 * jj + 1 < nsl;
            uint4 q;
            uint32_t spins = 0;
            long long t0 = 0;
            while (true) {
                q = ld_pair2(src);
                if (q.y == tag && (!two || q.w == tag)) break;
                if (((++spins) & 63u) == 0 && check_abort(tag, t0)) {
                    dead = true;
                    break;
                }
            }
            if (dead) break;
            const int e0 = rr * nsl + 2 * jj;
            const int m0 = h2map[e0];
            if (m0 >= 0) hs[m0] = __uint_as_float(q.x);
            if (two) {
                const int m1 = h2map[e0 + 1];
                if (m1 >= 0) hs[m1] = __uint_as_float(q.z);
            }
        }
        dead = __any_sync(0xffffffffu, dead);
        poller_sync();
        if (dead) return;
        if (p == 0 && pp.params_out != nullptr) {
            const int O = pl.O, T = pp.T;
            for (int i = pl_; i < O * BT; i += NPL) {
                const int o = i / BT, b = i % BT;
                if (b < pp.B) pp.params_out[((size_t)b * O + o) * T + t] = hs[i];
            }
            if (pl.head_kind == 2) {                         // the softmax sampler overwrites hs in place
                poller_sync();
                if (dead) return;
            }
        }
        for (int b = warp; b < BT; b += WN6_NPW) {
            sample_utt(t, b);
            if (t + 1 < pp.T) fetch_noise(t + 1, b);
        }
        poller_sync();
    }

    __device__ void poll_loop() {
        const int pl_ = warp * 32 + lane, NS = pl.NS, L = pl.L, T = pp.T;
        const int xin_floats = pl.xin_vals * BT;
        uint32_t n = 0;
        for (int t = 0; t < T && !dead; ++t) {
            for (int s = 0; s < NS; ++s, ++n) {
                const int par = n & 1;
                if (n >= 2) {
                    if (!wait_bar(&bar_free[par], ((n >> 1) - 1) & 1u, 0x10000000u | (uint32_t)s)) break;
                }
                float* xb = xin + (size_t)par * xin_floats;
                if (s == 0) {
                    if (t > 0) read_head_and_sample(t - 1, pl_);
                    if (dead) break;
                    write_x0(xb + pl.Ky * BT, pl_, true);
                    mbar_arrive(bar_x0);          // x_0 at the rows this block owns is in place (read by F1 in stage 1)
                } else {
                    int npairs;
                    if (s == 1) npairs = pl.Ky * BT;                       // x_0 is evaluated locally
                    else if (s <= L) npairs = (pl.Ky + pl.Kx) * BT;
                    else npairs = pl.Ksk * BT;
                    poll_pairs(pp.xbuf + wn6_ex_off(pl, s - 1, rank), npairs, n, xb, pl_);
                    if (dead) break;
                    if (s == 1) write_x0(xb + pl.Ky * BT, pl_, false);
                }
                mbar_arrive(&bar_in[par]);
            }
        }
        if (!dead) read_head_and_sample(T - 1, pl_);
    }

    // ======================================================================================
    // compute warps
    // ======================================================================================
    int rs_slot = 0;
    uint32_t rs_par = 0;
    __device__ __forceinline__ const float* acquire_blob(int t, int i) {
        if (i < pl.nres) {
            if (t == 0) wait_bar(&bar_full[i], 0, 0x80000000u | (uint32_t)i);
            return slots + (size_t)i * pl.slot_floats;
        }
        const int slot = pl.nres + rs_slot;
        wait_bar(&bar_full[slot], rs_par, 0x80000000u | (uint32_t)i);
        return slots + (size_t)slot * pl.slot_floats;
    }
    __device__ __forceinline__ void release_blob(int i) {
        if (i >= pl.nres) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[rs_slot]);
            if (++rs_slot == pl.nring) {
                rs_slot = 0;
                rs_par ^= 1u;
            }
        }
    }

    // one pass: two row quads (lane groups) x nit k-steps, then the partial sums go to the row owners
    __device__ __forceinline__ void run_pass(const Wn6Pass& ps, const float* __restrict__ blob, const float* __restrict__ xb,
                                             int par, int dpar, int xpar) {
        const int sub = lane & 15, g = lane >> 4;
        const float4* __restrict__ w = reinterpret_cast<const float4*>(blob + ps.w_off) + lane;
        const float* __restrict__ x = xb + (size_t)(ps.x_off + sub) * BT;
        float acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.f;
        const int nit = ps.nit;
#pragma unroll 2
        for (int j = 0; j < nit; ++j) {
            const float4 w4 = w[j * 32];
            float xv[BT];
            if constexpr (BT == 1) {
                xv[0] = x[j * 16];
            } else if constexpr (BT == 2) {
                const float2 t2 = *reinterpret_cast<const float2*>(x + j * 16 * BT);
                xv[0] = t2.x; xv[1] = t2.y;
            } else {
#pragma unroll
                for (int h = 0; h < BT / 4; ++h) {
                    const float4 t4 = *reinterpret_cast<const float4*>(x + j * 16 * BT + 4 * h);
                    xv[4 * h] = t4.x; xv[4 * h + 1] = t4.y; xv[4 * h + 2] = t4.z; xv[4 * h + 3] = t4.w;
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                acc[0 * BT + b] = fmaf(w4.x, xv[b], acc[0 * BT + b]);
                acc[1 * BT + b] = fmaf(w4.y, xv[b], acc[1 * BT + b]);
                acc[2 * BT + b] = fmaf(w4.z, xv[b], acc[2 * BT + b]);
                acc[3 * BT + b] = fmaf(w4.w, xv[b], acc[3 * BT + b]);
            }
        }
        reduce16<NV>(acc, lane);
        const int owner = ps.owner[g];
        if (owner < 0) return;
        const int CS = pl.CS;
        constexpr int NVL = NV >= 16 ? NV / 16 : 1;
        constexpr int DIV = NV >= 16 ? 1 : 16 / NV;         // lanes holding the same value
        if ((sub & (DIV - 1)) != 0) return;
        const int v0 = (sub / DIV) * NVL;                    // value index = row_in_quad*BT + b
        const int row = ps.dst_row[g] + v0 / BT, b0 = v0 % BT;
        const int dsel = ps.dst;
        const int nrow = dsel == 0 ? pl.nrow_c : (dsel == 1 ? pl.nrow_d : pl.nrow_x);
        const uint32_t base = smem_u32(dsel == 0 ? part : (dsel == 1 ? dpart : partx));
        const int bpar = dsel == 0 ? par : (dsel == 1 ? dpar : xpar);
        const uint32_t off = (uint32_t)((((size_t)bpar * nrow + row) * CS + rank) * BT + b0) * 4u;
        const uint32_t ra = mapa(base + off, (uint32_t)owner);
        const uint32_t rb = mapa(smem_u32(dsel == 0 ? &bar_part[bpar] : (dsel == 1 ? &bar_dpart[bpar] : &bar_partx[bpar])),
                                 (uint32_t)owner);
        if constexpr (NVL == 1) st_async_f32(ra, acc[0], rb);
        else st_async_f32x2(ra, acc[0], acc[1], rb);
    }

    __device__ void comp_loop() {
        const int cw = warp - WN6_W_COMP, NS = pl.NS, L = pl.L, T = pp.T;
        const int xin_floats = pl.xin_vals * BT;
        const bool prof = (pp.prof != nullptr) && cw == 0 && lane == 0;
        long long pc[4] = {0, 0, 0, 0}, tc = 0;
#define WN6_TICK(i) if (prof) { const long long now_ = clock64(); pc[i] += now_ - tc; tc = now_; }
        uint32_t n = 0, nd = 0, nl = 0;
        const float* blob = nullptr;
        for (int t = 0; t < T && !dead; ++t) {
            if (prof) tc = clock64();
            for (int s = 0; s < NS; ++s, ++n) {
                const int par = n & 1, kind = wn6_kind(pl, s);
                if (s <= L) blob = acquire_blob(t, s);
                WN6_TICK(0);
                if (!wait_bar(&bar_in[par], (n >> 1) & 1u, 0x08000000u | (uint32_t)s)) break;
                WN6_TICK(1);
                const float* xb = xin + (size_t)par * xin_floats;
                const int begin = pl.pass_begin[kind][cw], cnt = pl.pass_count[kind][cw], crit = pl.pass_crit[kind][cw];
                const int dpar = nd & 1, xpar = nl & 1;
                for (int i = 0; i < cnt; ++i) {
                    if (i == crit && nd >= 2) {
                        // deferred partials reuse the owners' buffer of two deferred stages ago: every owner of the
                        // cluster must have consumed it (credits sent by the DF warps)
                        if (!wait_bar<true>(&bar_dfree[dpar], ((nd >> 1) - 1) & 1u, 0x00400000u | (uint32_t)s)) break;
                    }
                    run_pass(passes[begin + i], blob, xb, par, dpar, xpar);
                    if (i + 1 == crit) WN6_TICK(2);
                }
                if (dead) break;
                // a warp without deferred passes in this stage still has to consume the credit of its turn
                if (pl.rows_d[kind] > 0 && cnt == crit && nd >= 2) {
                    if (!wait_bar<true>(&bar_dfree[dpar], ((nd >> 1) - 1) & 1u, 0x00400000u | (uint32_t)s)) break;
                }
                if (pl.rows_d[kind] > 0) ++nd;
                if (kind == WN6_K_LAYER) ++nl;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_free[par]);
                if (s < L || s == NS - 1) release_blob(wn6_blob_of_stage(pl, s));
                WN6_TICK(3);
            }
        }
        if (prof) {
            for (int i = 0; i < 4; ++i) pp.prof[(size_t)p * 16 + 8 + i] = pc[i];
        }
#undef WN6_TICK
    }

    // ======================================================================================
    // finalisers
    // ======================================================================================
    // modules.py:154  tanh(a) * sigmoid(g) with a single division:
    //   (1 - e^{-2a}) / ((1 + e^{-2a}) (1 + e^{-g}));  |a| is clamped where tanh has saturated in fp32.
    __device__ __forceinline__ static float gate(float a, float g) {
        const float ac = fminf(fmaxf(a, -15.0f), 15.0f);
        const float ea = expf(-2.0f * ac), eg = expf(-g);
        return (1.0f - ea) / ((1.0f + ea) * (1.0f + eg));
    }
    __device__ __forceinline__ float psum(const float* __restrict__ P, int row, int b) const {
        const float* q = P + ((size_t)row * pl.CS) * BT + b;
        float s = q[0];
        for (int r = 1; r < pl.CS; ++r) s += q[(size_t)r * BT];
        return s;
    }
    __device__ __forceinline__ uint32_t tx_bytes_c(int kind) const { return (uint32_t)(pl.rows_c[kind] * pl.CS * BT * 4); }
    __device__ __forceinline__ uint32_t tx_bytes_d(int kind) const { return (uint32_t)(pl.rows_d[kind] * pl.CS * BT * 4); }

    __device__ __forceinline__ uint32_t tx_bytes_x() const { return (uint32_t)(pl.rows_x * pl.CS * BT * 4); }

    // F0: gate rows (stages 0..L-1), total skip (stage L), head rows (stages L+1, L+2)
    __device__ void f0_loop() {
        const int NS = pl.NS, L = pl.L, T = pp.T, my = pl.my, ms = pl.ms, mo = pl.mo, RA4 = 4 * pl.qA;
        const uint32_t total = (uint32_t)T * (uint32_t)NS;
        const int ND = pl.rows_d[WN6_K_TAIL] > 0 ? L : L - 1;  // deferred stages per step
        const bool prof = (pp.prof != nullptr) && lane == 0;
        long long pc[4] = {0, 0, 0, 0}, tc = 0;
#define WN6_TICK(i) if (prof) { const long long now_ = clock64(); pc[i] += now_ - tc; tc = now_; }
        uint32_t n = 0;
        for (int t = 0; t < T && !dead; ++t) {
            if (prof) tc = clock64();
            for (int s = 0; s < NS; ++s, ++n) {
                const int par = n & 1, kind = wn6_kind(pl, s);
                if (s == 0) {
                    if (!wait_bar(bar_pre, (uint32_t)t & 1u, 0x02000000u)) break;     // pre-sums of this step are built
                }
                if (!wait_bar(&bar_part[par], (n >> 1) & 1u, 0x04000000u | (uint32_t)s)) break;
                WN6_TICK(0);
                __syncwarp();
                if (lane == 0 && n + 2 < total) mbar_expect_tx(&bar_part[par], tx_bytes_c(wn6_kind(pl, (s + 2) % NS)));
                const float* P = part + (size_t)par * pl.nrow_c * pl.CS * BT;
                const uint32_t tag = n + 1u;
                const long long ex = wn6_ex_off(pl, s, rank);
                if (kind == WN6_K_FIRST || kind == WN6_K_LAYER) {
                    for (int j = lane; j < my * BT; j += 32) {
                        const int i = j / BT, b = j % BT;
                        const float a = psum(P, 2 * i, b) + pre[((size_t)s * RA4 + 2 * i) * BT + b];
                        const float g = psum(P, 2 * i + 1, b) + pre[((size_t)s * RA4 + 2 * i + 1) * BT + b];
                        publish(ex + (long long)(c * my + i) * BT + b, gate(a, g), tag);
                    }
                } else if (kind == WN6_K_TAIL) {
                    // skip rows of layers 0..L-2 were accumulated by DF: its deferred stage of layer L-2 must be done
                    if (L >= 2) wait_count(s_ddone, t * ND + (L - 1), 0x02000002u);
                    for (int j = lane; j < ms * BT; j += 32) {
                        const int i = j / BT, b = j % BT;
                        // (s_0 + ... + s_{L-2}) + s_{L-1}, * sqrt(1/L), first ReLU of the head (wavenet.py:312-315)
                        float tot = psum(P, i, b) + bias[pl.bo_sb + (L - 1) * 4 * pl.qS + i];
                        if (L >= 2) tot = skipacc[i * BT + b] + tot;
                        publish(ex + (long long)(c * ms + i) * BT + b, fmaxf(tot * pl.skip_scale, 0.f), tag);
                    }
                } else if (kind == WN6_K_HEAD1) {
                    for (int j = lane; j < ms * BT; j += 32) {
                        const int i = j / BT, b = j % BT;
                        publish(ex + (long long)(c * ms + i) * BT + b, fmaxf(psum(P, i, b) + bias[pl.bo_ha + i], 0.f), tag);
                    }
                } else {
                    for (int j = lane; j < mo * BT; j += 32) {
                        const int i = j / BT, b = j % BT;
                        publish(ex + (long long)(c * mo + i) * BT + b, psum(P, i, b) + bias[pl.bo_hb + i], tag);
                    }
                }
                WN6_TICK(1);
            }
        }
        if (prof) {
            for (int i = 0; i < 4; ++i) pp.prof[(size_t)p * 16 + i] = pc[i];
        }
#undef WN6_TICK
    }

    // F1: the residual stream of the layer stages, modules.py:160-162  x_s = (conv1x1_out(y_{s-1}) + x_{s-1}) * sqrt(0.5)
    __device__ void f1_loop() {
        const int NS = pl.NS, L = pl.L, T = pp.T, mx = pl.mx;
        const float RSQRT2 = 0.70710678118654752440f;         // math.sqrt(0.5), modules.py:162
        const uint32_t total = (uint32_t)T * (uint32_t)(L - 1);
        uint32_t nl = 0;
        for (int t = 0; t < T && !dead; ++t) {
            for (int s = 1; s < L; ++s, ++nl) {
                const int xpar = nl & 1;
                const uint32_t n = (uint32_t)t * (uint32_t)NS + (uint32_t)s;
                if (s == 1) {
                    // x_0 at the rows this block owns was written by the pollers in stage 0 of this step
                    if (!wait_bar(bar_x0, (uint32_t)t & 1u, 0x02000001u)) break;
                }
                if (!wait_bar(&bar_partx[xpar], (nl >> 1) & 1u, 0x04100000u | (uint32_t)s)) break;
                __syncwarp();
                if (lane == 0 && nl + 2 < total) mbar_expect_tx(&bar_partx[xpar], tx_bytes_x());
                const float* P = partx + (size_t)xpar * pl.nrow_x * pl.CS * BT;
                const long long ex = wn6_ex_off(pl, s, rank);
                const float* xprev = (s == 1) ? x0own : xown;
                for (int j = lane; j < mx * BT; j += 32) {
                    const int i = j / BT, b = j % BT;
                    const float o = psum(P, i, b) + bias[pl.bo_xb + s * 4 * pl.qB + i];
                    const float xv = (o + xprev[i * BT + b]) * RSQRT2;
                    publish(ex + (long long)(pl.Ky + c * mx + i) * BT + b, xv, n + 1u);
                    xown[i * BT + b] = xv;
                }
            }
        }
    }

    // Everything of z_l(t) that does not depend on step t's exchanges: (folded) bias + global conditioning +
    // local-conditioning projection + the queued products of the older taps.
    __device__ void build_pre(int t) {
        const int L = pl.L, RA4 = 4 * pl.qA, kw = pl.kw, n = L * RA4 * BT;
        if (pl.C > 0) wait_bar<true>(&bar_cfull[t & 1], (uint32_t)(t >> 1) & 1u, 0x01000000u);
        if (dead) return;
        const float* cd = cond + (size_t)(t & 1) * L * RA4 * BT;
        for (int i = lane; i < n; i += 32) {
            const int l = i / (RA4 * BT), rem = i % (RA4 * BT);
            float v = sb[i];
            if (pl.C > 0) v += cd[i];
            for (int k = 0; k < kw - 1; ++k) {
                const int e = (l * (kw - 1) + k) * 3;
                v += ring[((size_t)ringtab[e] + ringtab[e + 2]) * RA4 * BT + rem];
            }
            pre[i] = v;
        }
        __syncwarp();
        if (lane == 0) {
            if (pl.C > 0) mbar_arrive(&bar_cempty[t & 1]);
            mbar_arrive(bar_pre);
        }
    }

    __device__ void dfin_loop() {
        const int L = pl.L, T = pp.T, kw = pl.kw, my = pl.my, ms = pl.ms, RA4 = 4 * pl.qA, CS = pl.CS;
        const bool tail_def = pl.rows_d[WN6_K_TAIL] > 0;
        const int ND = tail_def ? L : L - 1;
        const uint32_t total = (uint32_t)T * (uint32_t)ND;
        build_pre(0);
        uint32_t nd = 0;
        for (int t = 0; t < T && !dead; ++t) {
            for (int s = 1; s <= L; ++s) {
                const int kind = s < L ? WN6_K_LAYER : WN6_K_TAIL;
                if (pl.rows_d[kind] == 0) continue;
                const int dpar = nd & 1;
                if (!wait_bar<true>(&bar_dpart[dpar], (nd >> 1) & 1u, 0x00800000u | (uint32_t)s)) break;
                __syncwarp();
                if (lane == 0 && nd + 2 < total) {
                    // kind of the deferred stage two ahead
                    const uint32_t q = (nd + 2) % (uint32_t)ND;     // index inside its step: stage q+1
                    mbar_expect_tx(&bar_dpart[dpar], tx_bytes_d((int)q + 1 < L ? WN6_K_LAYER : WN6_K_TAIL));
                }
                const float* P = dpart + (size_t)dpar * pl.nrow_d * CS * BT;
                const int layer = s - 1;
                // older-tap products of `layer` -> history ring (consumed at steps t+d, t+2d, conv.py:32-44)
                for (int f = lane; f < (kw - 1) * 2 * my * BT; f += 32) {
                    const int b = f % BT, dr = f / BT, tap = dr / (2 * my), rr = dr % (2 * my);
                    const int e = (layer * (kw - 1) + tap) * 3;
                    ring[((size_t)ringtab[e] + ringtab[e + 2]) * RA4 * BT + rr * BT + b] = psum(P, dr, b);
                }
                // skip rows, accumulated in layer order (wavenet.py:312)
                if (kind == WN6_K_LAYER) {
                    for (int f = lane; f < ms * BT; f += 32) {
                        const int i = f / BT, b = f % BT;
                        const float h = psum(P, 4 * pl.qD + i, b) + bias[pl.bo_sb + layer * 4 * pl.qS + i];
                        skipacc[f] = (layer == 0) ? h : skipacc[f] + h;
                    }
                }
                __threadfence_block();
                __syncwarp();
                // credit: this owner is done with dpart[dpar]; every block of the cluster may send into it again
                if (lane < CS) mbar_arrive_remote(mapa(smem_u32(&bar_dfree[dpar]), (uint32_t)lane));
                ++nd;
                if (lane == 0) *s_ddone = (int)nd;
            }
            if (dead) break;
            // advance the ring positions to (t+1) mod delay, then the pre-sums of step t+1
            for (int i = lane; i < L * (kw - 1); i += 32) {
                const int pos = ringtab[i * 3 + 2] + 1;
                ringtab[i * 3 + 2] = (pos == ringtab[i * 3 + 1]) ? 0 : pos;
            }
            __syncwarp();
            if (t + 1 < T) build_pre(t + 1);
        }
    }
};

// ------------------------------------------------------------------------------------------
// kernel entry
// ------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(WN6_NTHREADS, 1)
wn6_kernel(const __grid_constant__ Wn6Plan pl, const __grid_constant__ Wn6Ptrs pp) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Engine<BT> eng(pl, pp, smem_raw);
    const int tid = threadIdx.x, p = blockIdx.x, warp = tid >> 5;
    const int nslots = pl.nres + pl.nring;
    const int NC = pl.NC, CS = pl.CS, rank = eng.rank, c = eng.c, L = pl.L;
    if (tid == 0) {
        for (int i = 0; i < nslots; ++i) mbar_init(&eng.bar_full[i], 1);
        for (int i = 0; i < pl.nring; ++i) mbar_init(&eng.bar_empty[i], WN6_NCW);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&eng.bar_cfull[i], 1);
            mbar_init(&eng.bar_cempty[i], 1);
            mbar_init(&eng.bar_in[i], 32 * WN6_NPW);
            mbar_init(&eng.bar_free[i], WN6_NCW);
            mbar_init(&eng.bar_part[i], 1);
            mbar_init(&eng.bar_dpart[i], 1);
            mbar_init(&eng.bar_partx[i], 1);
            mbar_init(&eng.bar_dfree[i], (uint32_t)CS);
        }
        mbar_init(eng.bar_pre, 1);
        mbar_init(eng.bar_x0, 32 * WN6_NPW);
        mbar_init(eng.bar_ps, 32 * WN6_NPW);
        *eng.s_abort = 0;
        *eng.s_ddone = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // pass table
    {
        const int nw = pl.npass * (int)(sizeof(Wn6Pass) / 4);
        const int* src = reinterpret_cast<const int*>(pp.passes);
        int* dst = reinterpret_cast<int*>(eng.passes);
        for (int i = tid; i < nw; i += WN6_NTHREADS) dst[i] = src[i];
    }
    // zero the history (== the reference's zero-initialised queue, conv.py:35-36) and the scratch buffers
    const int RA4 = 4 * pl.qA;
    if (pl.ring_in_smem) {
        const size_t n = (size_t)pl.ring_pos_total * RA4 * BT;
        for (size_t i = tid; i < n; i += WN6_NTHREADS) eng.ring[i] = 0.f;
    }
    for (int i = tid; i < 2 * pl.xin_vals * BT; i += WN6_NTHREADS) eng.xin[i] = 0.f;
    for (int i = tid; i < 2 * pl.nrow_c * CS * BT; i += WN6_NTHREADS) eng.part[i] = 0.f;
    for (int i = tid; i < 2 * pl.nrow_d * CS * BT; i += WN6_NTHREADS) eng.dpart[i] = 0.f;
    for (int i = tid; i < 2 * pl.nrow_x * CS * BT; i += WN6_NTHREADS) eng.partx[i] = 0.f;
    for (int i = tid; i < 4 * pl.qS * BT; i += WN6_NTHREADS) eng.skipacc[i] = 0.f;
    for (int i = tid; i < 8 * pl.qB * BT; i += WN6_NTHREADS) eng.xown[i] = 0.f;
    for (int i = tid; i < pl.O * BT; i += WN6_NTHREADS) eng.hs[i] = 0.f;
    for (int i = tid; i < pl.L * (pl.kw - 1); i += WN6_NTHREADS) {
        eng.ringtab[i * 3] = pp.ringtab[i * 2];           // offset of the ring (in positions)
        eng.ringtab[i * 3 + 1] = pp.ringtab[i * 2 + 1];   // delay D
        eng.ringtab[i * 3 + 2] = 0;                       // t mod D
    }
    // biases of the rows this block owns
    {
        const float* src = pp.bpack + (size_t)p * pl.cta_b_floats;
        for (int i = tid; i < pl.cta_b_floats; i += WN6_NTHREADS) eng.bias[i] = src[i];
    }
    // x_0 coefficient tables: K-slice (Kx entries) then own rows (4qB entries); [w or index | b]
    {
        const int Kx = pl.Kx, nown = 4 * pl.qB;
        for (int i = tid; i < Kx + nown; i += WN6_NTHREADS) {
            int g;
            float* tab;
            int k, kk;
            if (i < Kx) {
                g = wn6_slice_index(pl.R, NC, CS, pl.mx, rank, i);
                tab = eng.x0w; k = i; kk = Kx;
            } else {
                int base, cnt;
                wn6_own(pl.R, NC, CS, c, rank, base, cnt);
                k = i - Kx;
                g = k < cnt ? base + k : -1;
                tab = eng.x0w + 2 * Kx; kk = nown;
            }
            if (pl.input_kind == 0) tab[k] = g >= 0 ? pp.first_w[g] : 0.f;
            else tab[k] = __int_as_float(g);
            tab[kk + k] = g >= 0 ? pp.first_b[g] : 0.f;
        }
    }
    // map of the head-2 exchange (all ranks) to hs[o][b]
    {
        const int nsl = pl.Kh2 * BT;
        for (int i = tid; i < CS * nsl; i += WN6_NTHREADS) {
            const int rr = i / nsl, e = i % nsl, k = e / BT, b = e % BT;
            const int o = wn6_slice_index(pl.O, NC, CS, pl.mo, rr, k);
            eng.h2map[i] = o >= 0 ? o * BT + b : -1;
        }
    }
    {
        // static part of the pre-activation: (folded) conv bias + global-conditioning projection
        // (modules.py:148-152 recomputes Wg.g every step although g is constant; fold it once)
        int y0, ny;
        wn6_own(pl.G2, NC, CS, c, rank, y0, ny);
        const float* bsrc = pp.bpack + (size_t)p * pl.cta_b_floats + pl.bo_zb;
        const int n = L * RA4 * BT;
        for (int i = tid; i < n; i += WN6_NTHREADS) {
            const int b = i % BT, rr = (i / BT) % RA4, l = i / (BT * RA4);
            float v = 0.f;
            if ((rr >> 1) < ny) {
                v = bsrc[l * RA4 + rr];
                if (pp.gbias != nullptr && b < pp.B) {
                    const int grow = (rr & 1) ? pl.G2 + y0 + (rr >> 1) : y0 + (rr >> 1);
                    v += pp.gbias[((size_t)b * L + l) * pl.G + grow];
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
        for (int i = tid; i < BT * pl.O; i += WN6_NTHREADS) {
            const int b = i / pl.O, o = i % pl.O;
            eng.s_dense[i] = (dsrc && b < pp.B) ? dsrc[(size_t)b * stride + o] : 0.f;
        }
    }
    if (warp < WN6_NPW) {
        for (int b = warp; b < BT; b += WN6_NPW) eng.fetch_noise(0, b);
    }
    __syncthreads();
    if (tid == 0) {
        // arm the partial-sum barriers of the first two stages / deferred stages
        mbar_expect_tx(&eng.bar_part[0], eng.tx_bytes_c(wn6_kind(pl, 0)));
        mbar_expect_tx(&eng.bar_part[1], eng.tx_bytes_c(wn6_kind(pl, 1)));
        if (L >= 2) {
            mbar_expect_tx(&eng.bar_partx[0], eng.tx_bytes_x());
            mbar_expect_tx(&eng.bar_partx[1], eng.tx_bytes_x());
        }
        const int ND = pl.rows_d[WN6_K_TAIL] > 0 ? L : L - 1;
        if (ND > 0) {
            mbar_expect_tx(&eng.bar_dpart[0], eng.tx_bytes_d(1 < L ? WN6_K_LAYER : WN6_K_TAIL));
            const int q1 = 1 % ND;
            mbar_expect_tx(&eng.bar_dpart[1], eng.tx_bytes_d(q1 + 1 < L ? WN6_K_LAYER : WN6_K_TAIL));
        }
    }
    cluster_sync_all();      // every block of the cluster has initialised and armed its barriers

    if (warp < WN6_NPW) eng.poll_loop();
    else if (warp < WN6_W_F0) eng.comp_loop();
    else if (warp == WN6_W_F0) eng.f0_loop();
    else if (warp == WN6_W_F1) eng.f1_loop();
    else if (warp == WN6_W_DF) eng.dfin_loop();
    else if (warp == WN6_W_TMA) eng.tma_loop();
    else if (pl.C > 0) eng.cond_loop();

    // no block may exit while a peer can still write into its shared memory
    cluster_sync_all();
}

// gbias[b][l][row] = Wg_l[row,:] . g_b   (modules.py:148-152), once per call
__global__ void wn6_gbias_kernel(const float* __restrict__ wg, const float* __restrict__ g, float* __restrict__ out,
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
__global__ void wn6_sample_kernel(const float* __restrict__ y, int B, int O, int T, const float* __restrict__ u1,
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

}  // namespace wn6
