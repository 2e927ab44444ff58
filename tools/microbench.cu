// Micro-benchmarks that size the K1 epilogue on B200: TMEM read bandwidth, integer min / IMAD issue rates,
// tcgen05.mma.kind::i8 rate.  Build: make microbench ; run on the GPU box, prints cycles.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../monocularsfm_b200/csrc/ptx.cuh"
using namespace msfm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

// ---- 1. TMEM -> RF bandwidth: nwarps warps each load 32x32b.x32 `iters` times
__global__ void k_tmem_ld(int iters, long long* cyc, int* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<512>(&tptr);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(base + ((i * 32) & 511 & ~31u) % 480, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 8) acc ^= v[k];
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tptr);
}

// same, two loads in flight before the wait
__global__ void k_tmem_ld2(int iters, long long* cyc, int* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<512>(&tptr);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 2) {
        uint32_t v[32], w[32];
        ptx::tmem_ld_32x32(base + 0, v);
        ptx::tmem_ld_32x32(base + 32, w);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 8) acc ^= v[k] ^ w[k];
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tptr);
}

// ---- 2. ALU rates. mode 0: IMNMX chain x8 independent; 1: VIMNMX3; 2: IMAD; 3: IMAD + VIMNMX3 (epilogue mix);
//        4: IMAD+VIMNMX3+LDS.128 broadcast (full epilogue inner loop without TMEM)
template <int MODE>
__global__ void k_alu(int iters, long long* cyc, int* sink, const int* in) {
    __shared__ int4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_int4(threadIdx.x, threadIdx.x * 3, threadIdx.x * 5, threadIdx.x * 7);
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = in[threadIdx.x + k * 32];
    int x = in[threadIdx.x], y = in[threadIdx.x + 1];
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = min(a[k], x + k); x ^= i;
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = __vimin3_s32(a[k], x + 0, y); x ^= i;
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = a[k] * (-512) + x;
        } else if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                int ka = a[k] * (-512) + x, kb = a[k + 1] * (-512) + y;
                a[k] = __vimin3_s32(a[k], ka, kb); a[k + 1] ^= ka;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k += 4) {
                int4 c = sm[(i + k) & 63];
                int ka = c.x - (a[k] << 9), kb = c.y - (a[k + 1] << 9), kc = c.z - (a[k + 2] << 9), kd = c.w - (a[k + 3] << 9);
                x = __vimin3_s32(x, ka, kb); x = __vimin3_s32(x, kc, kd);
                a[k] += i; a[k + 1] ^= i; a[k + 2] -= i; a[k + 3] += 3;
            }
        }
    }
    long long t1 = clock64();
    int s = x ^ y;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= a[k];
    if (s == 0x7654321) sink[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 2b. column-direction reduction candidate: per 32 registers: 32 x redux.max + lane select, then one shared atomicMax
__global__ void k_redux(int iters, long long* cyc, int* sink, const int* in) {
    __shared__ int colmax[256];
    if (threadIdx.x < 256) colmax[threadIdx.x] = 0;
    int v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = in[(threadIdx.x + k * 7) & 1023];
    const int lane = threadIdx.x & 31;
    __syncthreads();
    long long t0 = clock64();
    int keep = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int m = __reduce_max_sync(0xffffffffu, v[k] + i);
            if (lane == k) keep = m;
        }
        atomicMax(&colmax[(i * 32 + lane) & 255], keep);
    }
    long long t1 = clock64();
    if (keep == 0x7654321) sink[0] = keep + colmax[lane];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 3. MMA rate: one thread issues `iters` x (4 x tcgen05.mma 128x256x32 i8) on garbage smem, then commit + wait
__global__ void k_mma(int iters, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc<512>(&tptr);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x == 32) {
        const uint64_t ad = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
        const uint64_t bd = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + 16384));
        constexpr uint32_t idesc = ptx::make_idesc_u8(128, 256);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::mma_i8_ss(tptr + (i & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, k > 0);
        }
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        cyc[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tptr);
}

// ---- 4. issue patterns of the real K1 loop, tensor pipe only (no epilogue): per "tile" 2 sub-tiles x (4 data + 1 ext) MMAs
//        of shape 128 x N x 32, A from shared memory (TS=0) or tensor memory (TS=1); OVH=1 adds what the issuing thread
//        does per tile in the real kernel: 3 waits on already-completed mbarriers + 3 commits.
template <int N, int TS, int OVH>
__global__ void k_mma_pat(int iters, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar, done_bar[3];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (2 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        for (int i = 0; i < 3; ++i) ptx::mbar_init(&done_bar[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) ptx::tmem_alloc<512>(&tptr);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x == 0) for (int i = 0; i < 3; ++i) ptx::mbar_arrive(&done_bar[i]);   // phase 0 complete
    __syncthreads();
    if (threadIdx.x == 32) {
        const uint64_t ad = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
        const uint64_t bd = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + 32768));
        constexpr uint32_t idesc = ptx::make_idesc_u8(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // OVH bit 0: 3 waits, bit 1: fence, bit 2: 3 commits, bit 3: 1 wait, bit 4: 1 commit
            if (OVH & 1) { ptx::mbar_wait(&done_bar[0], 0); ptx::mbar_wait(&done_bar[1], 0); ptx::mbar_wait(&done_bar[2], 0); }
            if (OVH & 8) { ptx::mbar_wait(&done_bar[0], 0); }
            if (OVH & 2) { ptx::tc_fence_after(); }
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                const uint32_t d = tptr + (i & 1) * 2 * N * (N <= 96 ? 1 : 0) + sub * N * (N <= 128 ? 1 : 0);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (TS) ptx::mma_i8_ts(d, tptr + 448 + (k & 3) * 8, bd + 2 * (k & 3), idesc, k > 0);
                    else ptx::mma_i8_ss(d, ad + sub * 1024 + 2 * (k & 3), bd + 2 * (k & 3), idesc, k > 0);
                }
            }
            if (OVH & 4) { ptx::mma_commit(&bar); ptx::mma_commit(&bar); ptx::mma_commit(&bar); }
            if (OVH & 16) { ptx::mma_commit(&bar); }
        }
        if (!(OVH & 20)) ptx::mma_commit(&bar);
        // drain: wait for the last commit phase (phase parity unknown under OVH: just spin on time)
        long long t1 = clock64();
        while (clock64() - t1 < 20000) {}
        cyc[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tptr);
}

template <int N, int TS, int OVH>
static int run_pat(int G, long long* d_cyc, long long* h, const char* name) {
    if (cudaFuncSetAttribute(k_mma_pat<N, TS, OVH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000) != cudaSuccess) return 1;
    const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) k_mma_pat<N, TS, OVH><<<G, 64, 70000>>>(iters, d_cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("pattern %s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < G; ++i) s += h[i];
    const double c = s / G / iters;
    printf("mma pattern %-34s: %.1f issue-cyc per tile (2 x 5 MMAs 128x%dx32; tensor floor %d cyc)\n", name, c, N, 10 * N / 2);
    return 0;
}

static double avg(long long* h, int n) { double s = 0; for (int i = 0; i < n; ++i) s += h[i]; return s / n; }

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    const int G = p.multiProcessorCount;
    long long* d_cyc; int* d_sink; int* d_in;
    CK(cudaMalloc(&d_cyc, G * 8)); CK(cudaMalloc(&d_sink, 64)); CK(cudaMalloc(&d_in, 4096)); CK(cudaMemset(d_in, 1, 4096));
    long long* h = new long long[G];
    const int it = 4096;
    for (int nw : {4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) k_tmem_ld<<<G, nw * 32>>>(it, d_cyc, d_sink);
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost));
        double c = avg(h, G);
        printf("tmem_ld  32x32b.x32 warps=%2d: %.1f cyc/iter/warp  -> %.1f B/cyc/SM\n", nw, c / it, nw * 4096.0 * it / c);
        for (int rep = 0; rep < 2; ++rep) k_tmem_ld2<<<G, nw * 32>>>(it, d_cyc, d_sink);
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost));
        c = avg(h, G);
        printf("tmem_ld2 (2 in flight) warps=%2d: %.1f cyc/ld/warp  -> %.1f B/cyc/SM\n", nw, c / it, nw * 4096.0 * it / c);
    }
    const char* names[5] = {"IMNMX x8", "VIMNMX3 x8", "IMAD x8", "2IMAD+VIMNMX3 x4", "epilogue mix (8 elem: 2 LDS.128, 8 sub/shl, 4 VIMNMX3)"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int nw : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k_alu<0><<<G, nw * 32>>>(it, d_cyc, d_sink, d_in); break;
                    case 1: k_alu<1><<<G, nw * 32>>>(it, d_cyc, d_sink, d_in); break;
                    case 2: k_alu<2><<<G, nw * 32>>>(it, d_cyc, d_sink, d_in); break;
                    case 3: k_alu<3><<<G, nw * 32>>>(it, d_cyc, d_sink, d_in); break;
                    default: k_alu<4><<<G, nw * 32>>>(it, d_cyc, d_sink, d_in); break;
                }
            }
            CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost));
            double c = avg(h, G);
            printf("alu %-55s warps=%2d: %.2f cyc/iter (per SM: %.2f cyc per warp-iter)\n", names[mode], nw, c / it, c / it / nw);
        }
    }
    for (int nw : {4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) k_redux<<<G, nw * 32>>>(it, d_cyc, d_sink, d_in);
        CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost));
        double c = avg(h, G);
        printf("redux column-max (32 REDUX.MAX + select + 1 ATOMS.MAX per 32x32 chunk) warps=%2d: %.1f cyc per chunk per warp, %.1f per SM-chunk\n", nw, c / it, c / it / nw);
    }
    CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000));
    for (int rep = 0; rep < 2; ++rep) k_mma<<<G, 64, 60000>>>(2048, d_cyc);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost));
    printf("mma i8 128x256x128 (4 x K32): %.1f cyc per tile (all %d SMs busy)\n", avg(h, G) / 2048, G);
    k_mma<<<1, 64, 60000>>>(2048, d_cyc);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("mma i8 128x256x128 (4 x K32): %.1f cyc per tile (1 SM)\n", (double)h[0] / 2048);
    run_pat<128, 1, 0>(G, d_cyc, h, "TS N=128 plain");
    run_pat<128, 1, 1>(G, d_cyc, h, "TS N=128 +3 waits");
    run_pat<128, 1, 2>(G, d_cyc, h, "TS N=128 +fence");
    run_pat<128, 1, 4>(G, d_cyc, h, "TS N=128 +3 commits");
    run_pat<128, 1, 8>(G, d_cyc, h, "TS N=128 +1 wait");
    run_pat<128, 1, 16>(G, d_cyc, h, "TS N=128 +1 commit");
    run_pat<128, 1, 26>(G, d_cyc, h, "TS N=128 +1 wait+fence+1 commit");
    run_pat<128, 0, 26>(G, d_cyc, h, "SS N=128 +1 wait+fence+1 commit");
    run_pat<128, 1, 7>(G, d_cyc, h, "TS N=128 +3w+f+3c");
    run_pat<64, 1, 26>(G, d_cyc, h, "TS N=64 +1 wait+fence+1 commit");
    run_pat<96, 1, 26>(G, d_cyc, h, "TS N=96 +1 wait+fence+1 commit");
    run_pat<256, 0, 26>(G, d_cyc, h, "SS N=256 +1 wait+fence+1 commit");
    printf("MICROBENCH DONE\n");
    return 0;
}
