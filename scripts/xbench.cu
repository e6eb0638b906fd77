// xbench.cu — microbenchmark of the tagged-pair L2 broadcast used by the synthesis kernel.
// P blocks; each round every block publishes its K/P (value,tag) pairs and then waits until it
// has seen all K pairs of that round.  Reports cycles per round for several polling structures.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xbench scripts/xbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint4 ld_pair2(const uint2* p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_pair(uint2* p, uint32_t v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v), "r"(tag) : "memory");
}
__device__ int g_abort = 0;
#define WATCHDOG(t0_) if (clock64() - (t0_) > 400000000LL || *((volatile int*)&g_abort)) { g_abort = 1; break; }
__device__ __forceinline__ void spin(long long cycles) {
    if (cycles <= 0) return;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}

// mode 0: every thread polls its own elements (k = tid + j*NT), 8-byte loads
// mode 1: warp 0 polls everything with 16-byte loads, hands over through shared memory
// mode 2: warps 0-1 poll (half each), 16-byte loads
// mode 3: like 0 but with 16-byte loads (each thread owns 2 adjacent elements)
__device__ __forceinline__ long long pidx(long long lin, int xc, int xs) { return xc ? (lin / xc) * xs + (lin % xc) : lin; }
__global__ void __launch_bounds__(256, 1)
xbench(uint2* buf, int K, int nslots, int rounds, int mode, int crit_delay, int def_delay, long long* out, int xc,
       long long xs_) {
    __shared__ float sh[2048];
    __shared__ int flag;
    const int tid = threadIdx.x, p = blockIdx.x, P = gridDim.x;
    const int per = K / P, k0 = p * per;
    const int xs = (int)xs_;
    uint32_t acc = 1;
    __syncthreads();
    const long long t_start = clock64();
    for (int r = 0; r < rounds; ++r) {
        const uint32_t tag = (uint32_t)r + 1u;
        const long long sbase = (long long)(r % nslots) * K;
        spin(crit_delay);
        if (tid < per) st_pair(buf + pidx(sbase + k0 + tid, xc, xs), acc + tid, tag);
        spin(def_delay);
        uint32_t sum = 0;
        const long long tw = clock64();
        if (mode == 0) {
            while (true) {
                WATCHDOG(tw)
                uint32_t bad = 0; sum = 0;
                for (int k = tid; k < K; k += 256) { const uint2 v = ld_pair(buf + pidx(sbase + k, xc, xs)); bad |= v.y ^ tag; sum += v.x; }
                if (!bad) break;
            }
        } else if (mode == 3) {
            while (true) {
                WATCHDOG(tw)
                uint32_t bad = 0; sum = 0;
                for (int k = 2 * tid; k < K; k += 512) { const uint4 v = ld_pair2(buf + pidx(sbase + k, xc, xs)); bad |= (v.y ^ tag) | (v.w ^ tag); sum += v.x + v.z; }
                if (!bad) break;
            }
        } else {
            const int nw = (mode == 1) ? 1 : 2;
            if (tid < 32 * nw) {
                while (true) {
                    WATCHDOG(tw)
                    uint32_t bad = 0;
                    for (int k = 2 * tid; k < K; k += 64 * nw) {
                        const uint4 v = ld_pair2(buf + pidx(sbase + k, xc, xs));
                        bad |= (v.y ^ tag) | (v.w ^ tag);
                        sh[k] = __uint_as_float(v.x); sh[k + 1] = __uint_as_float(v.z);
                    }
                    if (!bad) break;
                }
            }
            __syncthreads();
            for (int k = tid; k < K; k += 256) sum += __float_as_uint(sh[k]);
            __syncthreads();
        }
        acc = acc * 1664525u + (sum & 1u) + 1013904223u;
    }
    const long long t_end = clock64();
    if (tid == 0) out[p] = t_end - t_start;
    if (tid == 0 && acc == 0x12345) flag = 1;
}

// single-thread store->load round trip through L2 (one block)
__global__ void rtt(uint2* buf, int rounds, long long* out) {
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        st_pair(buf + (r & 63) * 16, r, r + 1);
        const long long tw = clock64();
        while (ld_pair(buf + (r & 63) * 16).y != (uint32_t)(r + 1)) { WATCHDOG(tw) }
    }
    out[0] = clock64() - t0;
}
// dependent load chain (pure L2 hit latency)
__global__ void chase(const uint2* buf, int rounds, long long* out) {
    uint32_t idx = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) idx = ld_pair(buf + idx).x;
    out[0] = clock64() - t0 + (idx == 77777);
}

int main(int argc, char** argv) {
    setvbuf(stdout, NULL, _IONBF, 0);
    const int rounds = 4000, nslots = 27;
    uint2* buf; long long* out;
    const size_t nb = (size_t)64 * 64 * 1024 * sizeof(uint2);
    cudaMalloc(&buf, nb); cudaMalloc(&out, 256 * sizeof(long long));
    int dev = 0, clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    long long h[256];
    // latency references
    cudaMemset(buf, 0, nb);
    rtt<<<1, 1>>>(buf, 4000, out); cudaDeviceSynchronize(); printf("rtt done: %s\n", cudaGetErrorString(cudaGetLastError())); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("store->load round trip (1 thread): %.0f cycles\n", (double)h[0] / 4000);
    {
        std::vector<uint2> hb(4096);
        for (int i = 0; i < 4096; ++i) hb[i] = make_uint2((i * 97 + 13) % 4096, 0);
        cudaMemcpy(buf, hb.data(), hb.size() * sizeof(uint2), cudaMemcpyHostToDevice);
        chase<<<1, 1>>>(buf, 4000, out); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("dependent ld.relaxed.gpu chain: %.0f cycles/load\n", (double)h[0] / 4000);
    }
    const int lay[][2] = {{0, 0}, {32, 544}, {16, 272}, {16, 1040}, {4, 516}, {32, 2080}, {1, 33}};
    const int delays[][2] = {{0, 0}, {500, 1500}};
    for (int di = 0; di < 2; ++di)
        for (int K : {256, 768})
            for (int P : {64, 128})
                for (int li = 0; li < 7; ++li) {
                    int mode = 0, xc = lay[li][0]; long long xs = lay[li][1];
                    cudaMemset(buf, 0, nb);
                    void* args[] = {&buf, (void*)&K, (void*)&nslots, (void*)&rounds, &mode, (void*)&delays[di][0],
                                    (void*)&delays[di][1], &out, &xc, &xs};
                    cudaError_t e = cudaLaunchCooperativeKernel((void*)xbench, dim3(P), dim3(256), args, 0, 0);
                    cudaDeviceSynchronize();
                    cudaError_t e2 = cudaGetLastError();
                    if (e != cudaSuccess || e2 != cudaSuccess) { printf("launch failed %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
                    cudaMemcpy(h, out, P * sizeof(long long), cudaMemcpyDeviceToHost);
                    int ab = 0; cudaMemcpyFromSymbol(&ab, g_abort, sizeof(int));
                    if (ab) { printf("WATCHDOG fired: K=%d P=%d layout=(%d,%lld)\n", K, P, xc, xs); ab = 0; cudaMemcpyToSymbol(g_abort, &ab, sizeof(int)); continue; }
                    long long mx = 0; for (int i = 0; i < P; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("delay(%4d,%4d) K=%4d P=%3d layout chunk=%2d stride=%4lld pairs : %7.0f cycles/round (minus delays %6.0f)\n",
                           delays[di][0], delays[di][1], K, P, xc, xs, (double)mx / rounds, (double)mx / rounds - delays[di][0] - delays[di][1]);
                }
    return 0;
}
