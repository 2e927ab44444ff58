// Micro-benchmark of the K1 epilogue's integer work: the per-row maximum over 4 groups of 32 accumulator values held in
// 128 registers (all operands distinct registers, as after tcgen05.ld), with the candidate instructions:
//   0: VIMNMX3 (s32 3-input max)      1: VIMNMX (s32 2-input max)
//   2: FMNMX3 (max.f32 d,a,b,c on the raw bit patterns; non-negative s32 < 2^31-2^23 order like their float images)
//   3: FMNMX  (2-input float max on the bit patterns)
// Prints cycles per "tile" (128 values per thread) with 8 warps per SM (2 per sub-partition, like the K1 epilogue), and
// checks that the float forms return bit-exact maxima on denormal-range patterns.
// Build: make microbench_alu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ int fmax3_bits(int a, int b, int c) {
    int d;
    asm("max.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int fmax2_bits(int a, int b) {
    int d;
    asm("max.f32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// MODE 0: VIMNMX3 two chains per group | 1: FMNMX3 two chains | 2: alternating 2-input FMNMX / VIMNMX (ptxas cannot fuse
// them into a 3-input op) four chains | 3: VIMNMX3 four chains | 4: no max work (baseline: register refill only)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_mnmx(int iters, const int* __restrict__ in, long long* cyc, int* out) {
    __shared__ int4 sm[2048];                    // 32 KB of "accumulator" values, re-read every iteration
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        sm[i] = make_int4(in[(4 * i) & 4095], in[(4 * i + 1) & 4095], in[(4 * i + 2) & 4095], in[(4 * i + 3) & 4095]);
    __syncthreads();
    int acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        int v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int4 x = sm[(threadIdx.x + (i + c * 8 + q) * 37) & 2047];
                v[c][4 * q] = x.x; v[c][4 * q + 1] = x.y; v[c][4 * q + 2] = x.z; v[c][4 * q + 3] = x.w;
            }
        const int z = (MODE == 0 || MODE == 3) ? static_cast<int>(0x80000000) : 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int m;
            if (MODE == 0) {
                int m0 = __vimax3_s32(z, v[c][0], v[c][1]), m1 = __vimax3_s32(z, v[c][2], v[c][3]);
#pragma unroll
                for (int e = 4; e < 32; e += 4) { m0 = __vimax3_s32(m0, v[c][e], v[c][e + 1]); m1 = __vimax3_s32(m1, v[c][e + 2], v[c][e + 3]); }
                m = max(m0, m1);
            } else if (MODE == 1) {
                int m0 = fmax3_bits(z, v[c][0], v[c][1]), m1 = fmax3_bits(z, v[c][2], v[c][3]);
#pragma unroll
                for (int e = 4; e < 32; e += 4) { m0 = fmax3_bits(m0, v[c][e], v[c][e + 1]); m1 = fmax3_bits(m1, v[c][e + 2], v[c][e + 3]); }
                m = max(m0, m1);
            } else if (MODE == 2) {
                int m0 = v[c][0], m1 = v[c][1], m2 = v[c][2], m3 = v[c][3];
#pragma unroll
                for (int e = 4; e < 32; e += 8) {
                    m0 = fmax2_bits(m0, v[c][e]); m1 = fmax2_bits(m1, v[c][e + 1]); m2 = fmax2_bits(m2, v[c][e + 2]); m3 = fmax2_bits(m3, v[c][e + 3]);
                    if (e + 4 < 32) { m0 = max(m0, v[c][e + 4]); m1 = max(m1, v[c][e + 5]); m2 = max(m2, v[c][e + 6]); m3 = max(m3, v[c][e + 7]); }
                }
                m = fmax2_bits(max(m0, m1), max(m2, m3));
            } else if (MODE == 3) {
                int m0 = __vimax3_s32(z, v[c][0], v[c][1]), m1 = __vimax3_s32(z, v[c][2], v[c][3]);
                int m2 = __vimax3_s32(z, v[c][4], v[c][5]), m3 = __vimax3_s32(z, v[c][6], v[c][7]);
#pragma unroll
                for (int e = 8; e < 32; e += 8) {
                    m0 = __vimax3_s32(m0, v[c][e], v[c][e + 1]); m1 = __vimax3_s32(m1, v[c][e + 2], v[c][e + 3]);
                    m2 = __vimax3_s32(m2, v[c][e + 4], v[c][e + 5]); m3 = __vimax3_s32(m3, v[c][e + 6], v[c][e + 7]);
                }
                m = max(__vimax3_s32(m0, m1, m2), m3);
            } else {
                m = v[c][0] ^ v[c][5] ^ v[c][10] ^ v[c][15] ^ v[c][16] ^ v[c][21] ^ v[c][26] ^ v[c][31];
            }
            acc += m;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// correctness of the float forms on integer bit patterns (denormal and low-normal range)
__global__ void k_check(const int* __restrict__ in, int n, int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 2 >= n) return;
    const int a = in[i], b = in[i + 1], c = in[i + 2];
    if (fmax3_bits(a, b, c) != max(a, max(b, c))) atomicAdd(bad, 1);
    if (fmax2_bits(a, b) != max(a, b)) atomicAdd(bad + 1, 1);
}

template <int MODE>
static int run(int G, const int* d_in, long long* d_cyc, int* d_out, const char* name) {
    const int iters = 2000;
    cudaMemset(d_cyc, 0, G * 8);
    for (int rep = 0; rep < 2; ++rep) k_mnmx<MODE><<<G, 256>>>(iters, d_in, d_cyc, d_out);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return 1; }
    static long long h[256];
    cudaMemcpy(h, d_cyc, G * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < G; ++i) s += h[i];
    printf("epilogue max over 4 x 32 registers, 8 warps/SM, %-28s: %7.1f cyc per tile\n", name, s / G / iters);
    return 0;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int G = p.multiProcessorCount;
    int *d_in, *d_out, *d_bad; long long* d_cyc;
    const int n = 1 << 20;
    CK(cudaMalloc(&d_in, n * 4)); CK(cudaMalloc(&d_out, G * 256 * 4)); CK(cudaMalloc(&d_cyc, G * 8)); CK(cudaMalloc(&d_bad, 8));
    int* h = new int[n];
    uint32_t s = 12345;
    for (int i = 0; i < n; ++i) {      // values as in K1: 0 .. ~2^23.4, plus edge patterns
        s = s * 1664525u + 1013904223u;
        const uint32_t r = s >> 8;
        h[i] = (i % 7 == 0) ? int(r & 0xff) : (i % 11 == 0) ? int(0x007fffff + (r & 3)) : int(r % 10400000u);
    }
    CK(cudaMemcpy(d_in, h, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, 8));
    k_check<<<(n + 255) / 256, 256>>>(d_in, n, d_bad);
    int bad[2]; CK(cudaMemcpy(bad, d_bad, 8, cudaMemcpyDeviceToHost));
    printf("float-max on integer bit patterns: %d mismatches (3-input), %d (2-input) over %d triples\n", bad[0], bad[1], n - 2);
    run<4>(G, d_in, d_cyc, d_out, "baseline (refill only)");
    run<0>(G, d_in, d_cyc, d_out, "VIMNMX3, 2 chains/group");
    run<3>(G, d_in, d_cyc, d_out, "VIMNMX3, 4 chains/group");
    run<1>(G, d_in, d_cyc, d_out, "FMNMX3, 2 chains/group");
    run<2>(G, d_in, d_cyc, d_out, "FMNMX/VIMNMX 2-input mix");
    printf("MICROBENCH ALU DONE\n");
    return 0;
}
