// xbench_v7.cu — microbenchmark of the single-hop exchange with separate polling warps (kernel v7 skeleton).
//
// P = 128 blocks.  A stage is a GEMV over a 768-value vector produced by the previous stage, 6 values per
// block.  Per stage a block
//   1. polls ALL 768 tagged (value, tag) pairs in L2 with NPOLL dedicated polling warps (16-byte loads = 2 pairs;
//      768 pairs = 384 loads -> 1 load per lane with 12 warps) into shared memory, then arrives on an mbarrier;
//   2. four compute warps (one row pair each: two 768-long gate rows or two 256-long residual rows) wake on that
//      mbarrier, read the vector from shared memory, FMA, warp butterfly, and lane 0 applies the gate and publishes
//      its value(s) straight to L2: no cross-warp reduction, no second hop.
// Reports cycles per stage and its split: exchange (publish -> all values visible and the compute warp awake) vs local
// chain (wake -> publish), and the visibility latency of a block's own value.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/xbench_v7 scripts/xbench_v7.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ld_pair2(const uint2* p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_pair(uint2* p, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ int g_abort = 0;
#define WATCHDOG(t0_) if (clock64() - (t0_) > 400000000LL || *((volatile int*)&g_abort)) { g_abort = 1; break; }

#define KTOT 768
#define KY 256
#define NSLOT 27
#define NCOMP 4

// pair index of element k of a stage slot: `spread` = 0 contiguous, else 4 pairs (one 32-byte sector) per 256-byte
// granule so that the 192 sectors of a stage land on as many L2 slices as possible
__device__ __forceinline__ long long pidx(int k, int spread) {
    if (spread == 1) return (long long)(k >> 2) * 32 + (k & 3);        // 4 pairs (32 B) per 256-byte granule
    if (spread == 2) return (long long)(k >> 1) * 16 + (k & 1);        // 2 pairs (16 B) per 128-byte line
    if (spread == 3) return (long long)(k >> 3) * 32 + (k & 7);        // 8 pairs (64 B) per 256-byte granule
    return (long long)k;
}

template <int NPOLL>
__global__ void __launch_bounds__(32 * (NPOLL + NCOMP), 1)
xv7(uint2* buf, long long slot_pairs, int rounds, int spread, int gate_delay, int backoff_ns, int lane_arrive, long long* out, float* check) {
    constexpr int NPL = 32 * NPOLL;           // polling lanes
    constexpr int NLD = KTOT / 2;             // 16-byte loads per stage
    __shared__ __align__(16) float xin[2][KTOT];
    __shared__ __align__(16) float wsm[NCOMP][2][KTOT];      // [warp][row][k]
    __shared__ __align__(8) uint64_t bar_in[2], bar_free[2];
    __shared__ long long t_pub_s, t_wake_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, p = blockIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_in[i], lane_arrive ? NPL : NPOLL); mbar_init(&bar_free[i], NCOMP); }
        t_pub_s = 0; t_wake_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NCOMP * 2 * KTOT; i += blockDim.x) (&wsm[0][0][0])[i] = 1.0f;
    __syncthreads();
    long long acc_exch = 0, acc_local = 0, acc_own = 0, acc_poll1 = 0;
    const long long t_start = clock64();
    float v_final = 0.f;
    if (warp < NPOLL) {
        const int pl = warp * 32 + lane;
        long long t_prev = clock64();
        for (int r = 0; r < rounds; ++r) {
            const int par = r & 1, use = r >> 1;
            if (use > 0) { const long long tw = clock64(); while (!mbar_try_wait(&bar_free[par], (use - 1) & 1)) { WATCHDOG(tw) } }
            // gate: the next vector cannot be complete earlier than `gate_delay` cycles after the previous one was
            if (gate_delay > 0 && r > 0) { while (clock64() - t_prev < gate_delay) {} }
            const long long tw = clock64();
            for (int j = pl; j < NLD; j += NPL) {
                float v0, v1;
                if (r == 0) { v0 = v1 = 1.0f; }
                else {
                    const uint32_t tag = (uint32_t)r;
                    const uint2* src = buf + (long long)((r - 1) % NSLOT) * slot_pairs + pidx(2 * j, spread);
                    uint4 q;
                    int tries = 0;
                    while (true) { q = ld_pair2(src); ++tries; if (q.y == tag && q.w == tag) break; if (backoff_ns > 0) __nanosleep(backoff_ns); WATCHDOG(tw) }
                    if (j == 6 * p / 2) { acc_own += clock64() - t_pub_s; acc_poll1 += tries; }
                    v0 = __uint_as_float(q.x); v1 = __uint_as_float(q.z);
                }
                *reinterpret_cast<float2*>(&xin[par][2 * j]) = make_float2(v0, v1);
            }
            t_prev = clock64();
            if (lane_arrive) mbar_arrive(&bar_in[par]);
            else { __syncwarp(); if (lane == 0) mbar_arrive(&bar_in[par]); }
        }
    } else {
        const int cw = warp - NPOLL;
        const bool gate_rows = cw < 2;
        const int K = gate_rows ? KTOT : KY;
        for (int r = 0; r < rounds; ++r) {
            const int par = r & 1, use = r >> 1;
            // weights of this stage are known before its input: keep them in registers
            float4 w0[KTOT / 128], w1[KTOT / 128];
#pragma unroll
            for (int j = 0; j < KTOT / 128; ++j) {
                if (j * 128 < K) {
                    w0[j] = *reinterpret_cast<const float4*>(&wsm[cw][0][(j * 32 + lane) * 4]);
                    w1[j] = *reinterpret_cast<const float4*>(&wsm[cw][1][(j * 32 + lane) * 4]);
                }
            }
            { const long long tw = clock64(); while (!mbar_try_wait(&bar_in[par], use & 1)) { WATCHDOG(tw) } }
            const long long t_wake = clock64();
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int j = 0; j < KTOT / 128; ++j) {
                if (j * 128 < K) {
                    const float4 x = *reinterpret_cast<const float4*>(&xin[par][(j * 32 + lane) * 4]);
                    a0 = fmaf(w0[j].x, x.x, a0); a1 = fmaf(w0[j].y, x.y, a1); a0 = fmaf(w0[j].z, x.z, a0); a1 = fmaf(w0[j].w, x.w, a1);
                    b0 = fmaf(w1[j].x, x.x, b0); b1 = fmaf(w1[j].y, x.y, b1); b0 = fmaf(w1[j].z, x.z, b0); b1 = fmaf(w1[j].w, x.w, b1);
                }
            }
            float a = a0 + a1, b = b0 + b1;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off);
                b += __shfl_xor_sync(0xffffffffu, b, off);
            }
            if (lane == 0) mbar_arrive(&bar_free[par]);
            uint2* dst = buf + (long long)(r % NSLOT) * slot_pairs;
            const uint32_t tag = (uint32_t)(r + 1);
            if (gate_rows) {
                if (lane == 0) {
                    // stand-in for the gate: two exponentials and a division that leave the value unchanged
                    const float ea = expf(-2.0f * fminf(a, 15.f)), eg = expf(-b);
                    const float gz = (1.0f - ea) / ((1.0f + ea) * (1.0f + eg));
                    const float v = a / (float)KTOT + 1.0f + 0.0f * gz;
                    st_pair(dst + pidx(6 * p + cw, spread), v, tag);
                    v_final = v;
                }
            } else {
                if (lane < 2) {
                    const float v = (lane == 0 ? a : b) / (float)KY + 1.0f;
                    st_pair(dst + pidx(6 * p + 2 + 2 * (cw - 2) + lane, spread), v, tag);
                }
            }
            if (cw == 0 && lane == 0) {
                const long long t_pub = clock64();
                if (r > 0) acc_exch += t_wake - t_pub_s;
                acc_local += t_pub - t_wake;
                t_pub_s = t_pub;
            }
        }
    }
    const long long t_end = clock64();
    __syncthreads();
    if (tid == 32 * NPOLL) { out[p * 8 + 0] = t_end - t_start; out[p * 8 + 1] = acc_exch; out[p * 8 + 2] = acc_local; check[p] = v_final; }
    if (warp < NPOLL && acc_poll1 > 0) { out[p * 8 + 3] = acc_own; out[p * 8 + 4] = acc_poll1; }
}

template <int NPOLL>
static void run(uint2* buf, long long slot_pairs, long long* out, float* check, int rounds, int spread, int gate_delay, int backoff_ns, int lane_arrive) {
    cudaMemset(buf, 0, (size_t)NSLOT * slot_pairs * sizeof(uint2));
    cudaMemset(out, 0, 128 * 8 * sizeof(long long));
    void* args[] = {&buf, &slot_pairs, &rounds, &spread, &gate_delay, &backoff_ns, &lane_arrive, &out, &check};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)xv7<NPOLL>, dim3(128), dim3(32 * (NPOLL + NCOMP)), args, 0, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("NPOLL=%d: launch failed %s / %s\n", NPOLL, cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); return; }
    long long h[128 * 8]; float hc[128];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, check, sizeof(hc), cudaMemcpyDeviceToHost);
    int ab = 0; cudaMemcpyFromSymbol(&ab, g_abort, sizeof(int));
    if (ab) { printf("NPOLL=%d spread=%d: WATCHDOG fired\n", NPOLL, spread); ab = 0; cudaMemcpyToSymbol(g_abort, &ab, sizeof(int)); return; }
    long long mx = 0; double se = 0, sl = 0, so = 0, st = 0; int bad = 0;
    for (int i = 0; i < 128; ++i) {
        mx = h[i * 8] > mx ? h[i * 8] : mx; se += h[i * 8 + 1]; sl += h[i * 8 + 2]; so += h[i * 8 + 3]; st += h[i * 8 + 4];
        if (hc[i] != (float)(rounds + 1)) ++bad;
    }
    printf("poll warps=%2d (%.1f loads/lane) spread=%d gate=%4d backoff=%3d lane_arrive=%d : %6.0f cycles/stage | exchange (publish->awake) %6.0f  local (awake->publish) %5.0f | own value visible after %6.0f cycles, %.2f polls  wrong=%d\n",
           NPOLL, 384.0 / (32 * NPOLL), spread, gate_delay, backoff_ns, lane_arrive, (double)mx / rounds, se / 128 / rounds, sl / 128 / rounds, so / 128 / rounds, st / 128 / rounds, bad);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    uint2* buf; long long* out; float* check;
    const long long slot_pairs = 192LL * 32 + 64;      // room for the spread layout
    cudaMalloc(&buf, (size_t)NSLOT * slot_pairs * sizeof(uint2));
    cudaMalloc(&out, 128 * 8 * sizeof(long long));
    cudaMalloc(&check, 128 * sizeof(float));
    const int rounds = 6000;
    // 1. arrive per lane vs per warp
    for (int la : {1, 0}) { run<12>(buf, slot_pairs, out, check, rounds, 1, 0, 0, la); run<4>(buf, slot_pairs, out, check, rounds, 1, 0, 0, la); }
    // 2. layouts
    for (int sp : {0, 1, 2, 3}) { run<12>(buf, slot_pairs, out, check, rounds, sp, 0, 0, 0); run<4>(buf, slot_pairs, out, check, rounds, sp, 0, 0, 0); }
    // 3. gating delay and back-off between attempts
    for (int sp : {1, 3})
        for (int gd : {0, 800, 1200, 1600, 2000, 2400})
            for (int bo : {0, 100}) {
                run<12>(buf, slot_pairs, out, check, rounds, sp, gd, bo, 0);
                run<8>(buf, slot_pairs, out, check, rounds, sp, gd, bo, 0);
                run<4>(buf, slot_pairs, out, check, rounds, sp, gd, bo, 0);
            }
    return 0;
}
