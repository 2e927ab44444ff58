// Micro-benchmark of the fp64 building blocks of the band Cholesky (csrc/ba_band.cu): what ONE warp (or one thread of it) pays
// per dependent operation — the solver's spine is a chain of such operations, so these latencies, not the fp64 rate, set
// its pace.  Prints cycles per operation (clock64 around 4096 dependent / independent operations, one CTA of 32 threads, then
// the same with 8 warps to see contention).
//   dfma_dep        dependent DFMA chain (latency)
//   dfma_ind8       8 independent DFMA chains (issue rate of one warp)
//   dmul_shfl_dfma  the back substitution's step: multiply -> 64-bit shuffle -> FMA
//   rsqrt_lib       rsqrt(double) chain            rsqrt_f32seed   MUFU.RSQ seed + one Newton step chain
//   rcp_f32seed     MUFU.RCP seed + one Newton step chain
//   lds_dep         dependent shared-memory loads (pointer chase)
//   one_thread_*    the same chains executed by lane 0 only (the 12 x 12 diagonal sub-block is factored by one thread)
// Build: make microbench_f64 ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kN = 4096;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double seed, long long* cyc, double* out, int one_thread) {
    __shared__ int chase[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    double a = seed + threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + i;
    int idx = threadIdx.x & 1023;
    const bool active = !one_thread || (threadIdx.x & 31) == 0;
    const long long t0 = clock64();
    if (active) {
        if (MODE == 0) {
#pragma unroll 16
            for (int i = 0; i < kN; ++i) a = fma(a, b, c);
        } else if (MODE == 1) {
#pragma unroll 2
            for (int i = 0; i < kN / 8; ++i) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = fma(v[q], b, c);
            }
        } else if (MODE == 2) {
#pragma unroll 8
            for (int i = 0; i < kN; ++i) {
                const double x = __shfl_sync(0xffffffffu, a * b, i & 31);
                a = fma(-x, c, a);
            }
        } else if (MODE == 3) {
#pragma unroll 4
            for (int i = 0; i < kN; ++i) a = rsqrt(a) + 1.5;
        } else if (MODE == 4) {
#pragma unroll 4
            for (int i = 0; i < kN; ++i) {
                float s;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(__double2float_rn(a)));
                const double y = static_cast<double>(s);
                const double e = fma(-(0.5 * a) * y, y, 0.5);
                a = fma(y, e, y) + 1.5;
            }
        } else if (MODE == 5) {
#pragma unroll 4
            for (int i = 0; i < kN; ++i) {
                float s;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(__double2float_rn(a)));
                const double r = static_cast<double>(s);
                a = fma(r, fma(-a, r, 1.0), r) + 1.5;
            }
        } else if (MODE == 6) {
#pragma unroll 8
            for (int i = 0; i < kN; ++i) idx = chase[idx];
        } else if (MODE == 7) {
            // 47 independent FMAs behind one dependent one: the right-looking row solve's shape
            double x[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = seed + i;
#pragma unroll 1
            for (int i = 0; i < kN / 12; ++i) {
#pragma unroll
                for (int kk = 0; kk < 11; ++kk) {
#pragma unroll
                    for (int j = kk + 1; j < 12; ++j) x[j] = fma(-x[kk], c, x[j]);
                }
                x[0] = x[11] * b;
            }
            a = x[0] + x[5];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double r = a + idx;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
static int run(const char* name, int ops) {
    long long* cyc; double* out;
    CK(cudaMalloc(&cyc, 8 * sizeof(long long)));
    CK(cudaMalloc(&out, 8 * 256 * sizeof(double)));
    for (int one = 0; one < 2; ++one)
        for (int threads : {32, 256}) {
            k<MODE><<<1, threads>>>(1.25, cyc, out, one);
            CK(cudaDeviceSynchronize());
            k<MODE><<<1, threads>>>(1.25, cyc, out, one);
            CK(cudaDeviceSynchronize());
            long long h = 0;
            CK(cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost));
            printf("%-16s %s %d warp(s): %8.1f cycles per step (%d steps)\n", name, one ? "lane 0 only," : "full warps, ", threads / 32,
                   static_cast<double>(h) / ops, ops);
        }
    cudaFree(cyc); cudaFree(out);
    return 0;
}

int main() {
    if (run<0>("dfma_dep", kN)) return 1;
    if (run<1>("dfma_ind8", kN)) return 1;
    if (run<2>("dmul_shfl_dfma", kN)) return 1;
    if (run<3>("rsqrt_lib", kN)) return 1;
    if (run<4>("rsqrt_f32seed", kN)) return 1;
    if (run<5>("rcp_f32seed", kN)) return 1;
    if (run<6>("lds_dep", kN)) return 1;
    if (run<7>("tri12_fma", (kN / 12) * 67)) return 1;
    return 0;
}
