// xbench_cluster.cu — microbenchmark of the 2-D (cluster x rank) exchange that kernel v6 is built on.
//
// P = NC*CS blocks in NC thread-block clusters of CS.  A "stage" is a GEMV whose K input values were
// produced by the previous stage, one value group per block.  Cluster c owns a block of output rows;
// rank r of every cluster owns the K-slice made of the values finalised by the rank-r blocks of ALL
// clusters.  Per stage a block therefore
//   1. polls only its K/CS slice of the tagged (value, tag) pairs in L2 (each pair is polled by NC
//      blocks instead of P),
//   2. multiplies it with its [cluster rows x K-slice] weight tile from shared memory,
//   3. sends the partial sums to the owner block of every row through DSMEM (st.async + complete_tx),
//   4. the owner sums CS partials, applies the gate and publishes its values for the next stage.
// Reports cycles per stage.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xbench_cluster scripts/xbench_cluster.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__device__ int g_abort = 0;
#define WATCHDOG(t0_) if (clock64() - (t0_) > 400000000LL || *((volatile int*)&g_abort)) { g_abort = 1; break; }

// shapes of "config 2": stage vector 768 values, 256 gate pairs + 512 residual rows over P=128 blocks
#define KTOT 768
#define NPOLLW 2        // polling warps
#define NCOMPW 4        // compute warps
#define NTH (32 * (NPOLLW + NCOMPW + 1))
#define NSLOT 27

template <int CS>
__global__ void __launch_bounds__(NTH, 1)
xcluster(uint2* buf, int rounds, int extra_spin, long long* out, float* check) {
    constexpr int NC = 128 / CS;
    constexpr int KS = KTOT / CS;             // K-slice of this rank
    constexpr int VPB = KTOT / 128;           // values finalised per block (6)
    constexpr int ROWS = VPB * CS;            // rows the cluster's blocks send partials for, per block role
    constexpr int TPR = (ROWS > 64) ? 1 : ((ROWS > 32) ? 2 : ((ROWS > 16) ? 4 : ((ROWS > 8) ? 8 : 16)));
    static_assert(ROWS * TPR <= 32 * NCOMPW, "rows x threads-per-row must fit the compute warps");
    constexpr int NIT = KS / (4 * TPR) > 0 ? KS / (4 * TPR) : 1;
    __shared__ __align__(16) float xin[2][KS];
    __shared__ __align__(16) float part[2][VPB][CS];
    __shared__ __align__(16) float wsm[NIT][32 * NCOMPW][4];
    __shared__ __align__(8) uint64_t bar_in[2], bar_free[2], bar_part[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank(), c = blockIdx.x / CS;
    constexpr int NPL = KS / 2;               // 16-byte loads per stage (2 pairs each)
    constexpr int NACT = NPL < 32 * NPOLLW ? NPL : 32 * NPOLLW;   // polling lanes
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_in[i], NACT);
            mbar_init(&bar_free[i], NCOMPW);
            mbar_init(&bar_part[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NIT * 32 * NCOMPW * 4; i += NTH) (&wsm[0][0][0])[i] = 1.0f;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar_part[0], VPB * CS * 4);
        mbar_expect_tx(&bar_part[1], VPB * CS * 4);
    }
    cluster_sync();
    long long t_poll = 0, t_comp = 0, t_fin = 0;
    const long long t_start = clock64();
    float v_final = 0.f;
    if (warp < NPOLLW) {
        // ---------------- pollers
        const int pl = warp * 32 + lane;      // polling lane index
        if (pl < NACT) {
            for (int r = 0; r < rounds; ++r) {
                const int par = r & 1, use = r >> 1;
                if (use > 0) { const long long tw = clock64(); while (!mbar_try_wait(&bar_free[par], (use - 1) & 1)) { WATCHDOG(tw) } }
                const long long tw = clock64();
                for (int j = pl; j < NPL; j += NACT) {
                    float v0, v1;
                    if (r == 0) { v0 = v1 = 1.0f; }
                    else {
                        const uint32_t tag = (uint32_t)r;
                        const uint2* src = buf + ((size_t)((r - 1) % NSLOT) * CS + rank) * KS + 2 * j;
                        uint4 q;
                        while (true) { q = ld_pair2(src); if (q.y == tag && q.w == tag) break; WATCHDOG(tw) }
                        v0 = __uint_as_float(q.x); v1 = __uint_as_float(q.z);
                    }
                    *reinterpret_cast<float2*>(&xin[par][2 * j]) = make_float2(v0, v1);
                }
                if (pl == 0) t_poll += clock64() - tw;
                mbar_arrive(&bar_in[par]);
            }
        }
    } else if (warp < NPOLLW + NCOMPW) {
        // ---------------- compute: ROWS rows x KS, TPR threads per row
        const int ct = tid - 32 * NPOLLW;
        const int row = ct / TPR, sub = ct % TPR;
        const uint32_t part_base = smem_u32(&part[0][0][0]), bar_base = smem_u32(&bar_part[0]);
        for (int r = 0; r < rounds; ++r) {
            const int par = r & 1, use = r >> 1;
            { const long long tw = clock64(); while (!mbar_try_wait(&bar_in[par], use & 1)) { WATCHDOG(tw) } }
            const long long tc = clock64();
            float acc = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const float4 w = *reinterpret_cast<const float4*>(&wsm[it][ct][0]);
                const int k = ((it * TPR + sub) * 4) % KS;
                const float4 x = *reinterpret_cast<const float4*>(&xin[par][k]);
                acc = fmaf(w.x, x.x, acc); acc1 = fmaf(w.y, x.y, acc1); acc2 = fmaf(w.z, x.z, acc2); acc3 = fmaf(w.w, x.w, acc3);
            }
            acc = (acc + acc1) + (acc2 + acc3);
#pragma unroll
            for (int off = TPR / 2; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_free[par]);
            if (sub == 0 && row < ROWS) {
                const int owner = row / VPB, i = row % VPB;
                const uint32_t ra = mapa(part_base + (uint32_t)(((par * VPB + i) * CS + rank) * 4), (uint32_t)owner);
                const uint32_t rb = mapa(bar_base + (uint32_t)(par * 8), (uint32_t)owner);
                st_async_f32(ra, acc, rb);
            }
            if (ct == 0) t_comp += clock64() - tc;
            if (extra_spin > 0) { const long long t0 = clock64(); while (clock64() - t0 < extra_spin) {} }
        }
    } else {
        // ---------------- finaliser
        for (int r = 0; r < rounds; ++r) {
            const int par = r & 1, use = r >> 1;
            { const long long tw = clock64(); while (!mbar_try_wait(&bar_part[par], use & 1)) { WATCHDOG(tw) } }
            const long long tc = clock64();
            float s = 0.f;
            if (lane < VPB) {
#pragma unroll
                for (int j = 0; j < CS; ++j) s += part[par][lane][j];
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&bar_part[par], VPB * CS * 4);   // arm the next use of this buffer
            if (lane < VPB) {
                // stand-in for the gate: two exponentials and a division that leave the value unchanged
                const float ea = expf(-2.0f * fminf(s, 15.f)), eg = expf(-s);
                const float gz = (1.0f - ea) / ((1.0f + ea) * (1.0f + eg));
                const float v = s / (float)KTOT + 1.0f + 0.0f * gz;
                st_pair(buf + ((size_t)(r % NSLOT) * CS + rank) * KS + c * VPB + lane, v, (uint32_t)(r + 1));
                v_final = v;
            }
            if (lane == 0) t_fin += clock64() - tc;
        }
    }
    const long long t_end = clock64();
    cluster_sync();
    if (tid == 32 * (NPOLLW + NCOMPW)) { out[blockIdx.x * 4 + 0] = t_end - t_start; out[blockIdx.x * 4 + 3] = t_fin; check[blockIdx.x] = v_final; }
    if (tid == 0) out[blockIdx.x * 4 + 1] = t_poll;
    if (tid == 32 * NPOLLW) out[blockIdx.x * 4 + 2] = t_comp;
}

template <int CS>
static void run(uint2* buf, long long* out, float* check, int rounds, int spin, bool coop) {
    cudaMemset(buf, 0, (size_t)NSLOT * KTOT * sizeof(uint2));
    cudaMemset(out, 0, 128 * 4 * sizeof(long long));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(128); lc.blockDim = dim3(NTH); lc.dynamicSmemBytes = 0; lc.stream = 0;
    cudaLaunchAttribute la[2]; int na = 0;
    la[na].id = cudaLaunchAttributeClusterDimension; la[na].val.clusterDim.x = CS; la[na].val.clusterDim.y = 1; la[na].val.clusterDim.z = 1; ++na;
    if (coop) { la[na].id = cudaLaunchAttributeCooperative; la[na].val.cooperative = 1; ++na; }
    lc.attrs = la; lc.numAttrs = na;
    if (CS > 8) cudaFuncSetAttribute(xcluster<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int ncl = -1;
    cudaOccupancyMaxActiveClusters(&ncl, xcluster<CS>, &lc);
    cudaError_t e = cudaLaunchKernelEx(&lc, xcluster<CS>, buf, rounds, spin, out, check);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        printf("CS=%2d coop=%d: launch failed: %s / %s (max active clusters %d)\n", CS, (int)coop, cudaGetErrorString(e), cudaGetErrorString(e2), ncl);
        cudaGetLastError();
        return;
    }
    long long h[128 * 4]; float hc[128];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, check, sizeof(hc), cudaMemcpyDeviceToHost);
    int ab = 0; cudaMemcpyFromSymbol(&ab, g_abort, sizeof(int));
    if (ab) { printf("CS=%2d: WATCHDOG fired\n", CS); ab = 0; cudaMemcpyToSymbol(g_abort, &ab, sizeof(int)); return; }
    long long mx = 0; double sp = 0, sc = 0, sf = 0; int bad = 0;
    for (int i = 0; i < 128; ++i) {
        mx = h[i * 4] > mx ? h[i * 4] : mx; sp += h[i * 4 + 1]; sc += h[i * 4 + 2]; sf += h[i * 4 + 3];
        if (hc[i] != (float)(rounds + 1)) ++bad;
    }
    printf("CS=%2d NC=%3d coop=%d spin=%4d max_clusters=%3d : %7.0f cycles/stage   poll %6.0f  compute+send %5.0f  finalise+publish %5.0f   wrong=%d (v=%g)\n",
           CS, 128 / CS, (int)coop, spin, ncl, (double)mx / rounds, sp / 128 / rounds, sc / 128 / rounds, sf / 128 / rounds, bad, hc[0]);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    uint2* buf; long long* out; float* check;
    cudaMalloc(&buf, (size_t)NSLOT * KTOT * sizeof(uint2) + 4096);
    cudaMalloc(&out, 128 * 4 * sizeof(long long));
    cudaMalloc(&check, 128 * sizeof(float));
    const int rounds = 20000;
    for (int rep = 0; rep < 2; ++rep)
        for (int spin : {0, 300}) {
            run<1>(buf, out, check, rounds, spin, true);
            run<2>(buf, out, check, rounds, spin, true);
            run<4>(buf, out, check, rounds, spin, true);
            run<8>(buf, out, check, rounds, spin, true);
            run<8>(buf, out, check, rounds, spin, false);
            run<16>(buf, out, check, rounds, spin, true);
            run<16>(buf, out, check, rounds, spin, false);
        }
    return 0;
}
