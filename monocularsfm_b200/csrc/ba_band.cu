// Band Cholesky of the reduced camera system, hand-written: ONE cooperative kernel factors the renumbered system
// S = L L^T inside its band, carries the right-hand sides through the forward substitution, and back-substitutes.
//
// Replaces the linear solver Ceres runs behind ceres::Solve for DENSE_SCHUR / SPARSE_SCHUR
// (src/Optimizer/CeresBundleOptimizer.cpp:264-273,293).  Round 2 first walked the band as a block-tridiagonal chain of library
// calls (cuSOLVER potrf + cuBLAS trsm / syrk per super-block, ba_solver.cu): at BASELINE configs[4] (7968 unknowns, band of
// 126 cameras) that chain is latency-bound — 334 us per potrf(756) panel kernel, 1439 small trsm kernels, 9.2 ms per solve
// (profiles/r02_ba_launches_library_chain.csv) — no faster than the dense potrf it replaced.  Here the whole factorisation is
// one launch: the lower band is stored as 48 x 48 fp64 tiles (8 cameras per tile row), and per tile column j
//     D   one CTA:      L_jj = chol(T_jj)            blocked 4 x 12; one thread factors each 12 x 12 diagonal sub-block in registers
//     P   nbk CTAs:     L_Ij = T_Ij L_jj^-T          thread = row of the tile, forward substitution against L_jj in shared memory
//                       y_j  = L_jj^-1 y_j           (the right-hand sides ride along as one more panel task)
//     U   all CTAs:     T_IK -= L_Ij L_Kj^T          one 48x48x48 tile product per CTA (3x3 register blocks), j < K <= I <= j + nbk
//                       y_I  -= L_Ij y_j
// separated by grid barriers on a monotonic counter (cooperative launch: every CTA is resident); D of column j+1 runs on CTA 0
// right behind its update of that tile, under the other CTAs' updates (two barriers per column).  Then CTA 0 back-substitutes.
// fp64 throughout.  Tensor cores are not used (north_star: not for the sparse BA).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ba_types.cuh"
#include "ptx.cuh"
#include "launch_count.hpp"

namespace msfm {
namespace band {

constexpr int NB = 48;             // tile order: 8 cameras
constexpr int kThreads = 256;
constexpr int kLd = NB + 1;        // shared-memory leading dimension (doubles)

struct Params {
    double* tiles;     // [R][nbk + 1][NB][NB] row-major tiles of the lower band: slot d of block row I holds tile (I, I - nbk + d)
    double* y;         // [nrhs][R * NB] right-hand sides / solutions (renumbered order)
    double* dinv;      // [R * NB] reciprocals of the diagonal of L (the substitutions multiply instead of dividing)
    unsigned int* bar; // grid barrier counter (zeroed before the launch)
    int32_t* info;     // != 0: a pivot was not positive
    double* zfar;      // [nrhs][R * NB] back substitution: sums of the tiles further than 4 from the diagonal, added by the helper CTAs
    unsigned int* flags;   // [nrhs][R + 1]: done[col] = helper sums that have reached zfar[col]; [R] = columns solved so far
    long long* dbg;    // optional [16]: cycles CTA 0 spent in D, P, U, the grid barriers, the back substitution (+ three parts of it) (MSFM_BAND_DEBUG=1)
    int32_t R, nbk, nrhs;
};

__device__ __forceinline__ double* tile_ptr(const Params& p, int I, int J) {
    return p.tiles + (static_cast<size_t>(I) * (p.nbk + 1) + (J - I + p.nbk)) * (NB * NB);
}

// global tile (row-major) <-> shared [NB][kLd]
__device__ __forceinline__ void load_tile(const double* __restrict__ g, double* s) {
    for (int i = threadIdx.x; i < NB * NB; i += kThreads) s[(i / NB) * kLd + (i % NB)] = g[i];
}

// ---- D: Cholesky of a 48 x 48 tile in shared memory sA (lower triangle used), L written to sA and to global g.
// Thread r < 48 owns row r in registers; after step k every thread has published L[r][k], so row k (needed by all at the next
// dot product) is complete in shared memory.
// ---- D: Cholesky of a 48 x 48 tile in shared memory, blocked 4 x 12: ONE THREAD factors each 12 x 12 diagonal sub-block in its
// registers (the 48 reciprocal square roots are the sequential spine of the whole solver; nothing else shares their critical
// path), then a thread per row solves the sub-panel and all threads update the trailing part.  12 barriers per tile instead
// of the 96 of the first version (thread = row, one column published per pair of barriers: 556 cycles per pivot) — and the
// version that kept the tile in one warp's registers and moved columns by shuffles was slower still (in-order issue: every
// shuffle -> FMA pair stalls, 1080 cycles per pivot); profiles/r02_band_cholesky_phases.txt.
// sA: the tile ([NB][kLd], lower triangle used) -> L; sInv: reciprocals of the diagonal of L.  All threads of the CTA call.
constexpr int kSub = 12;
__device__ __noinline__ void diag_factor(double* sA, double* sInv, double* __restrict__ g, double* __restrict__ dinv, int32_t* info) {
    for (int b = 0; b < NB; b += kSub) {
        if (threadIdx.x == 0) {
            double m[kSub][kSub];
#pragma unroll
            for (int i = 0; i < kSub; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) m[i][j] = sA[(b + i) * kLd + b + j];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < kSub; ++k) {
                double pv = m[k][k];
                if (!(pv > 0.0)) { bad = true; pv = 1.0; }
                const double ri = rsqrt(pv);
                m[k][k] = pv * ri;
                sInv[b + k] = ri;
#pragma unroll
                for (int i = k + 1; i < kSub; ++i) m[i][k] *= ri;
#pragma unroll
                for (int i = k + 1; i < kSub; ++i)
#pragma unroll
                    for (int j = k + 1; j <= i; ++j) m[i][j] -= m[i][k] * m[j][k];
            }
            if (bad) atomicExch(info, 1);
#pragma unroll
            for (int i = 0; i < kSub; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) sA[(b + i) * kLd + b + j] = m[i][j];
        }
        __syncthreads();
        const int n = NB - b - kSub;                    // rows below the sub-block
        if (threadIdx.x < n) {
            double* row = sA + (b + kSub + threadIdx.x) * kLd + b;
            double x[kSub];
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                double v = row[j];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= x[k] * sA[(b + j) * kLd + b + k];
                x[j] = v * sInv[b + j];
            }
#pragma unroll
            for (int j = 0; j < kSub; ++j) row[j] = x[j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n * n; i += kThreads) {
            const int rr = i / n, cc = i - rr * n;
            if (cc > rr) continue;
            const double* pr = sA + (b + kSub + rr) * kLd + b;
            const double* pc = sA + (b + kSub + cc) * kLd + b;
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int k = 0; k < kSub; k += 2) { v0 += pr[k] * pc[k]; v1 += pr[k + 1] * pc[k + 1]; }
            sA[(b + kSub + rr) * kLd + b + kSub + cc] -= v0 + v1;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < NB * NB; i += kThreads) {
        const int r = i / NB, c = i - r * NB;
        g[i] = c <= r ? sA[r * kLd + c] : 0.0;
    }
    if (threadIdx.x < NB) dinv[threadIdx.x] = sInv[threadIdx.x];
}

// ---- P: rows x of a tile solved against L: x <- x L^-T, i.e. x[c] = (x[c] - sum_{k<c} x[k] L[c][k]) / L[c][c].
// One thread per row, the row in registers, no communication.  sLs holds L with row c SCALED by 1 / L[c][c] (load_tile_scaled),
// so that x[c] = x[c] / L[c][c] - sum_k x[k] Ls[c][k], and the loop runs column by column (right-looking): as soon as x[k] is
// final it is removed from all later unknowns, the first of which is the next one to become final — ONE dependent FMA per
// unknown (8.8 cycles, tools/microbench_f64.cu) with the other 46 - k FMAs of the column issued in its shadow.  The row-by-row
// form with four partial sums had the last FMA, two adds and the multiply by 1 / L[c][c] on the chain of every unknown.
__device__ __forceinline__ void solve_row(double x[NB], const double* sLs, const double* sInv) {
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] *= sInv[c];
#pragma unroll
    for (int k = 0; k < NB - 1; ++k) {
        const double xk = x[k];
#pragma unroll
        for (int c = k + 1; c < NB; ++c) x[c] = fma(-xk, sLs[c * kLd + k], x[c]);
    }
}
__device__ __forceinline__ void load_tile_scaled(const double* __restrict__ g, const double* __restrict__ dinv, double* s) {
    for (int i = threadIdx.x; i < NB * NB; i += kThreads) s[(i / NB) * kLd + (i % NB)] = g[i] * dinv[i / NB];
}

// ---- U: C -= A B^T for 48 x 48 tiles; A, B staged in shared memory TRANSPOSED ([k][row]); 16 x 16 threads, 3 x 3 outputs each
__device__ void tile_update(const double* __restrict__ gA, const double* __restrict__ gB, double* __restrict__ gC, double* sA, double* sB) {
    for (int i = threadIdx.x; i < NB * NB; i += kThreads) {
        const int row = i / NB, k = i % NB;
        sA[k * kLd + row] = gA[i];
        sB[k * kLd + row] = gB[i];
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
        const double a0 = sA[k * kLd + 3 * ty], a1 = sA[k * kLd + 3 * ty + 1], a2 = sA[k * kLd + 3 * ty + 2];
        const double b0 = sB[k * kLd + 3 * tx], b1 = sB[k * kLd + 3 * tx + 1], b2 = sB[k * kLd + 3 * tx + 2];
        acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2;
        acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2;
        acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) gC[(3 * ty + i) * NB + 3 * tx + j] -= acc[i][j];
    __syncthreads();
}

// Grid barrier on a monotonic counter (every CTA is resident: cooperative launch).  ~1.5 us instead of the ~5 us measured for
// cooperative_groups' grid.sync() on 137 CTAs (profiles/r02_band_cholesky_phases.txt).
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int& target, unsigned int nblk) {
    target += nblk;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile unsigned int*>(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}
// The two halves of the same barrier, for a CTA that has work between its arrival and the point where it needs the others
// (CTA 0: the diagonal chain).  target counts as in grid_barrier.
__device__ __forceinline__ void grid_arrive(unsigned int* ctr, unsigned int& target, unsigned int nblk) {
    target += nblk;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
    }
}
__device__ __forceinline__ void grid_wait(unsigned int* ctr, unsigned int target) {
    if (threadIdx.x == 0) {
        while (*reinterpret_cast<volatile unsigned int*>(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// ---- z += T^T x for one 48 x 48 row-major tile by one warp (back substitution).  The tile is walked flat, 32 consecutive
// doubles per step (72 steps), so every load instruction of the warp reads 256 contiguous bytes; lane l meets the columns
// l, (l + 32) % 48 and l + 16 — for steps 0, 1, 2 (mod 3) — and rows (32 i + l) / 48.  Lanes l and l + 16 hold the same three
// columns in rotated order: one exchange finishes the sums.  (The first layout, three adjacent columns per lane, touched 24
// sectors per load instruction instead of 8 and made the load/store unit the limit of the whole back substitution.)
__device__ __forceinline__ int flat_row(int i, int lane) { return (32 * i + lane) / NB; }
__device__ __forceinline__ void tile_fetch(const double* __restrict__ tile, int lane, double (&t)[72]) {
#pragma unroll
    for (int i = 0; i < 72; ++i) t[i] = tile[32 * i + lane];
}
// four partial sums per column: dependent chains of 6 FMAs instead of 24
__device__ __forceinline__ void tile_dot_regs(const double (&t)[72], const double* sx, int lane, double& a, double& b, double& c) {
    double pa[4] = {0.0, 0.0, 0.0, 0.0}, pb[4] = {0.0, 0.0, 0.0, 0.0}, pc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 72; i += 3) {
        pa[(i / 3) & 3] += t[i] * sx[flat_row(i, lane)];
        pb[(i / 3) & 3] += t[i + 1] * sx[flat_row(i + 1, lane)];
        pc[(i / 3) & 3] += t[i + 2] * sx[flat_row(i + 2, lane)];
    }
    a = (pa[0] + pa[1]) + (pa[2] + pa[3]);
    b = (pb[0] + pb[1]) + (pb[2] + pb[3]);
    c = (pc[0] + pc[1]) + (pc[2] + pc[3]);
}
// ---- the diagonal tile of the back substitution, solved four unknowns at a time.
// invert_diag_blocks: lane bk < 12 inverts the 4 x 4 diagonal block bk of L (row-major tile sL, reciprocals of the diagonal sI)
// into W[bk][16] (row-major, lower triangle).  Off the critical path: done by an idle warp a column ahead.
__device__ __forceinline__ void invert_diag_blocks(const double* sL, const double* sI, double* W, int lane) {
    if (lane >= NB / 4) return;
    const int u0 = 4 * lane;
    double w[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double di = sI[u0 + i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k > i) { w[i][k] = 0.0; continue; }
            if (k == i) { w[i][k] = di; continue; }
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m >= k && m < i) acc += sL[(u0 + i) * NB + u0 + m] * w[m][k];
            w[i][k] = -di * acc;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) W[lane * 16 + 4 * i + k] = w[i][k];
}
// solve_diag_blocked: L^T x = b by one warp; lane l holds unknowns l (b0) and l + 32 (b1, lanes < 16).  Per block of four
// unknowns, from the last: four shuffles fetch the block's b, every lane forms x_blk = W_blk^T b_blk (10 FMAs, three levels
// deep), the owners keep their x, the lanes below eliminate the four unknowns from their own (rows of L read ahead of the
// chain).  12 steps of shuffle + ~6 dependent operations instead of 48 steps of multiply -> shuffle -> FMA (55 cycles each,
// tools/microbench_f64.cu).
__device__ __forceinline__ void solve_diag_blocked(const double* sL, const double* W, int l, double& b0, double& b1) {
#pragma unroll
    for (int bk = NB / 4 - 1; bk >= 0; --bk) {
        const int u0 = 4 * bk;
        const bool hi = u0 >= 32;                                  // compile-time: the loop is unrolled
        double bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bb[i] = __shfl_sync(0xffffffffu, hi ? b1 : b0, (u0 + i) & 31);
        const double* w = W + bk * 16;
        const double x3 = w[15] * bb[3];
        const double x2 = w[10] * bb[2] + w[14] * bb[3];
        const double x1 = (w[5] * bb[1] + w[9] * bb[2]) + w[13] * bb[3];
        const double x0 = (w[0] * bb[0] + w[4] * bb[1]) + (w[8] * bb[2] + w[12] * bb[3]);
        // branch-free: every lane forms the update, selects decide who keeps what (a five-way divergent branch per block cost
        // more than the arithmetic)
        const double l0 = sL[(u0 + 0) * NB + l], l1 = sL[(u0 + 1) * NB + l], l2 = sL[(u0 + 2) * NB + l], l3 = sL[(u0 + 3) * NB + l];
        const double upd0 = (b0 - (l0 * x0 + l1 * x1)) - (l2 * x2 + l3 * x3);
        if (hi) {
            // unknowns 32 + l of lanes l < u0 - 32 are below the block; unknowns l (all lanes) too
            const int r = l - (u0 - 32);
            const int lc = 32 + (l & 15);                          // a valid column for every lane; lanes >= 16 hold no second unknown
            const double h0 = sL[(u0 + 0) * NB + lc], h1 = sL[(u0 + 1) * NB + lc], h2 = sL[(u0 + 2) * NB + lc], h3 = sL[(u0 + 3) * NB + lc];
            const double upd1 = (b1 - (h0 * x0 + h1 * x1)) - (h2 * x2 + h3 * x3);
            const double xs = r == 0 ? x0 : (r == 1 ? x1 : (r == 2 ? x2 : x3));
            b1 = r < 0 ? upd1 : (r < 4 ? xs : b1);
            b0 = upd0;
        } else {
            const int r = l - u0;
            const double xs = r == 0 ? x0 : (r == 1 ? x1 : (r == 2 ? x2 : x3));
            b0 = r < 0 ? upd0 : (r < 4 ? xs : b0);
        }
    }
}
// ---- flags between the CTAs of the back substitution (all resident: cooperative launch).  Waits are bounded: a protocol bug
// must become a trapped kernel, never a hang.
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_at_least(const unsigned int* p, unsigned int want) {
    if (ld_acquire(p) >= want) return;
    const long long t0 = clock64();
    while (ld_acquire(p) < want) {
        if (clock64() - t0 > 4000000000LL) {           // ~2 s
            printf("msfm: band back substitution flag wait timed out (block %d thread %d want %u have %u)\n", blockIdx.x, threadIdx.x, want, ld_acquire(p));
            __trap();
        }
    }
}
// lanes < 16 add the finished sums of columns l, l + 16, l + 32 to zc
__device__ __forceinline__ void tile_dot_commit(double a, double b, double c, int lane, double* zc) {
    const double a2 = __shfl_xor_sync(0xffffffffu, a, 16), b2 = __shfl_xor_sync(0xffffffffu, b, 16), c2 = __shfl_xor_sync(0xffffffffu, c, 16);
    if (lane < 16) {
        zc[lane] += a + b2;
        zc[lane + 16] += c + a2;
        zc[lane + 32] += b + c2;
    }
}

// dynamic shared memory, used by the back substitution only (CTA 0: a third diagonal tile buffer, ring of pending column
// sums, three sets of pivots / right-hand sides / inverted diagonal blocks, two of helper sums, three mbarriers; helpers: one
// copy of x_j per warp)
extern __shared__ __align__(128) unsigned char band_dyn[];
constexpr size_t kDynSmem = NB * NB * sizeof(double) + 16 * NB * sizeof(double) + 3 * (NB / 4) * 16 * sizeof(double) + 64;

__global__ void __launch_bounds__(kThreads, 1)
band_cholesky_kernel(Params p) {
    __shared__ __align__(16) double sA[NB * kLd];
    __shared__ __align__(16) double sB[NB * kLd];
    __shared__ double sv[6 * NB];
    __shared__ double sInv[NB];
    const int R = p.R, nbk = p.nbk, Npad = R * NB;
    const int bid = blockIdx.x, nblk = gridDim.x;
    unsigned int bar_target = 0;
    const bool timing = p.dbg != nullptr && bid == 0 && threadIdx.x == 0;
    long long tD = 0, tP = 0, tU = 0, tS = 0, t0 = 0;
#define BAND_TICK(acc) do { if (timing) { const long long t1 = clock64(); acc += t1 - t0; t0 = t1; } } while (0)
    if (timing) t0 = clock64();

    bool sA_has_L = false;                                  // CTA 0: sA / sInv still hold the factor of the current diagonal tile
    if (bid == 0) {
        load_tile(tile_ptr(p, 0, 0), sA);
        __syncthreads();
        diag_factor(sA, sInv, tile_ptr(p, 0, 0), p.dinv, p.info);
        __syncthreads();
        sA_has_L = true;
    }
    BAND_TICK(tD);
    grid_barrier(p.bar, bar_target, nblk);
    BAND_TICK(tS);
    for (int j = 0; j < R; ++j) {
        const int m = min(nbk, R - 1 - j);                 // tile rows below the diagonal in this column
        // ---- P: tasks 0 .. m-1 = tiles (j + 1 + t, j); task m = the right-hand sides
        // CTA 0's chain P(j+1, j) -> update of (j+1, j+1) -> D(j+1) is the critical path of the whole factorisation (every other
        // CTA waits for it at the next barrier), so it keeps its data on chip: L_jj is still in sA from D(j) (no reload), the next
        // diagonal tile is fetched into registers before the solve, the update reads the freshly solved panel tile from shared
        // memory and hands its result to D(j+1) in shared memory (round 2's first version went through global memory three
        // times here: ~4k of the 37k cycles of a column).  Only when the grid is at least as large as the task lists.
        const bool chain = bid == 0 && m > 0 && nblk > m * (m + 1) / 2;
        double cnext[3][3];                                          // thread (ty, tx): its 3 x 3 block of tile (j+1, j+1)
        for (int t = bid; t <= m; t += nblk) {
            if (chain && t == 0 && sA_has_L) {
                for (int i = threadIdx.x; i < NB * NB; i += kThreads) {
                    const int r = i / NB, c = i - r * NB;
                    sB[r * kLd + c] = sA[r * kLd + c] * sInv[r];
                }
                const double* gC = tile_ptr(p, j + 1, j + 1);
                const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) cnext[i][jj] = gC[(3 * ty + i) * NB + 3 * tx + jj];
                __syncthreads();                                     // sA (L_jj) has been read: the panel tile may replace it
            } else {
                load_tile_scaled(tile_ptr(p, j, j), p.dinv + static_cast<size_t>(j) * NB, sB);     // L_jj, row c scaled by 1 / L_cc
                if (threadIdx.x < NB) sInv[threadIdx.x] = p.dinv[static_cast<size_t>(j) * NB + threadIdx.x];
                if (chain && t == 0) {
                    const double* gC = tile_ptr(p, j + 1, j + 1);
                    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) cnext[i][jj] = gC[(3 * ty + i) * NB + 3 * tx + jj];
                }
            }
            if (t < m) {
                double* g = tile_ptr(p, j + 1 + t, j);
                // the tile through shared memory: coalesced global access, one row per thread afterwards
                load_tile(g, sA);
                __syncthreads();
                if (threadIdx.x < NB) {
                    double x[NB];
#pragma unroll
                    for (int c = 0; c < NB; ++c) x[c] = sA[threadIdx.x * kLd + c];
                    solve_row(x, sB, sInv);
#pragma unroll
                    for (int c = 0; c < NB; ++c) sA[threadIdx.x * kLd + c] = x[c];
                }
                __syncthreads();
                for (int i = threadIdx.x; i < NB * NB; i += kThreads) g[i] = sA[(i / NB) * kLd + (i % NB)];
            } else {
                __syncthreads();
                if (threadIdx.x < p.nrhs) {
                    double* yj = p.y + static_cast<size_t>(threadIdx.x) * Npad + static_cast<size_t>(j) * NB;
                    double x[NB];
#pragma unroll
                    for (int c = 0; c < NB; ++c) x[c] = yj[c];
                    solve_row(x, sB, sInv);
#pragma unroll
                    for (int c = 0; c < NB; ++c) yj[c] = x[c];
                }
            }
            __syncthreads();
        }
        BAND_TICK(tP);
        // ---- U: tasks 0 .. m(m+1)/2 - 1 = tiles (I, K), j < K <= I <= j + m; the last task = right-hand side updates.
        //      Task 0 is the next diagonal tile (j+1, j+1) -= L_{j+1,j} L_{j+1,j}^T: it needs only the panel tile CTA 0 has just
        //      solved itself, so CTA 0 arrives at the barrier between P and U without waiting, updates and factors the next
        //      diagonal tile (D of column j+1) while the other CTAs are still in P, the barrier and their updates, and only then
        //      waits for the others' panel tiles (for its further update tasks, when the grid is smaller than the task list).
        const int ntile = m * (m + 1) / 2;
        if (bid == 0) {
            grid_arrive(p.bar, bar_target, nblk);
            if (chain) {
                // sA still holds the panel tile X = L_{j+1,j} this CTA has just solved (row-major, stride kLd): C -= X X^T
                const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
                double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
                for (int k = 0; k < NB; ++k) {
                    const double a0 = sA[(3 * ty) * kLd + k], a1 = sA[(3 * ty + 1) * kLd + k], a2 = sA[(3 * ty + 2) * kLd + k];
                    const double b0 = sA[(3 * tx) * kLd + k], b1 = sA[(3 * tx + 1) * kLd + k], b2 = sA[(3 * tx + 2) * kLd + k];
                    acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2;
                    acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2;
                    acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2;
                }
                __syncthreads();                                     // every thread has read X
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) sA[(3 * ty + i) * kLd + 3 * tx + jj] = cnext[i][jj] - acc[i][jj];
                __syncthreads();
                BAND_TICK(tU);
                diag_factor(sA, sInv, tile_ptr(p, j + 1, j + 1), p.dinv + static_cast<size_t>(j + 1) * NB, p.info);
                __syncthreads();
                sA_has_L = true;
                BAND_TICK(tD);
            } else if (m > 0) {
                tile_update(tile_ptr(p, j + 1, j), tile_ptr(p, j + 1, j), tile_ptr(p, j + 1, j + 1), sA, sB);
                BAND_TICK(tU);
                load_tile(tile_ptr(p, j + 1, j + 1), sA);            // written by this CTA just above
                __syncthreads();
                diag_factor(sA, sInv, tile_ptr(p, j + 1, j + 1), p.dinv + static_cast<size_t>(j + 1) * NB, p.info);
                __syncthreads();
                sA_has_L = true;
                BAND_TICK(tD);
            }
            grid_wait(p.bar, bar_target);
        } else {
            grid_barrier(p.bar, bar_target, nblk);
        }
        BAND_TICK(tS);
        for (int t = bid == 0 ? nblk : bid; t <= ntile; t += nblk) {
            if (t < ntile) {
                int a = static_cast<int>((sqrtf(8.0f * static_cast<float>(t) + 1.0f) - 1.0f) * 0.5f);
                while (a * (a + 1) / 2 > t) --a;
                while ((a + 1) * (a + 2) / 2 <= t) ++a;
                const int b = t - a * (a + 1) / 2;           // 0 <= b <= a < m
                const int I = j + 1 + a, K = j + 1 + b;
                tile_update(tile_ptr(p, I, j), tile_ptr(p, K, j), tile_ptr(p, I, K), sA, sB);
                sA_has_L = false;
            } else if (m > 0) {
                // y_I -= L_Ij y_j for the m tile rows below: thread = (tile row a, row r), the 48 products of a row in flight together
                for (int q = 0; q < p.nrhs; ++q) {
                    const double* yj = p.y + static_cast<size_t>(q) * Npad + static_cast<size_t>(j) * NB;
                    if (threadIdx.x < NB) sv[threadIdx.x] = yj[threadIdx.x];
                    __syncthreads();
                    for (int i = threadIdx.x; i < m * NB; i += kThreads) {
                        const int a = i / NB, r = i - a * NB;
                        const double2* L2 = reinterpret_cast<const double2*>(tile_ptr(p, j + 1 + a, j) + r * NB);
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int c = 0; c < NB / 2; ++c) {
                            const double2 l = L2[c];
                            s0 += l.x * sv[2 * c];
                            s1 += l.y * sv[2 * c + 1];
                        }
                        p.y[static_cast<size_t>(q) * Npad + static_cast<size_t>(j + 1 + a) * NB + r] -= s0 + s1;
                    }
                    __syncthreads();
                }
            }
        }
        BAND_TICK(tU);
        grid_barrier(p.bar, bar_target, nblk);
        BAND_TICK(tS);
    }
    // ---- back substitution: x_j = L_jj^-T (y_j - z_j),  z_j = sum_{I > j} L_Ij^T x_I, columns from the last.
    // The chain x_{j+1} -> x_j is sequential, but only through the diagonal solve and the NEAREST tiles; everything else is
    // bandwidth: 17 tiles of 18 KB per column, and one SM draws ~42 B/clk from L2 (7.4k cycles per column — the single-CTA
    // versions sat at 9.3-10k whatever the prefetching, profiles/r02_band_cholesky_phases.txt).  So the tiles are spread:
    //   CTA 0  solves.  Warp 0: the 48 unknowns of a column, four at a time against inverted 4 x 4 diagonal blocks (the diagonal
    //          tile, its pivots and y_j arrive by bulk copy a column ahead; warp 1 inverts the blocks).  Warps 1..4: the tiles
    //          at distance 1..4 (tile (j, j - w)), loaded into registers BEFORE x_j exists, multiplied right after, summed into a
    //          shared-memory ring of pending column sums.  Warp 5 collects the helpers' sums of the NEXT column; warp 6
    //          publishes the previous x (global store + release of a monotonic counter) under the current solve; warp 7
    //          starts the bulk copies two columns ahead and inverts the diagonal blocks one column ahead.
    //   CTA h  (helper, one per distance d = 4 + h <= nbk): each of its warps takes every 8th column j, loads tile (j, j - d)
    //          ahead of time, waits for x_j, multiplies, adds into zfar[j - d] with fp64 reductions and releases done[j - d].
    //          A column's last helper sum comes from x_{j+5}: it has more than three column times to arrive.
    // Right-looking, so no tile ever waits for an x on the critical path except the four of CTA 0.
    {
        using namespace ptx;
        constexpr int kWarps = kThreads / 32;
        constexpr int kNear = 4;
        constexpr uint32_t kTileBytes = NB * NB * sizeof(double);
        // the warp index through a shuffle: the compiler then knows it is uniform across the warp and compiles the shuffles of the
        // warp-specialised branches below as plain SHFL; with `threadIdx.x >> 5` every one of them was wrapped in a convergence
        // barrier (WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE) and completed before the next could start — 96 of them, 5.1k cycles per
        // column in the diagonal solve alone
        const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
        const int n_helpers = max(nbk - kNear, 0);
        if (bid > n_helpers) return;
        double* dyn = reinterpret_cast<double*>(band_dyn);
        if (bid > 0) {
            // ------------------------------------------------------------------ helper of distance d
            const int d = kNear + bid;
            double* xs = dyn + warp * NB;                          // this warp's copy of x_j
            for (int q = 0; q < p.nrhs; ++q) {
                const double* y = p.y + static_cast<size_t>(q) * Npad;
                double* zfar = p.zfar + static_cast<size_t>(q) * Npad;
                unsigned int* flags = p.flags + static_cast<size_t>(q) * (R + 1);     // [0..R): done[col], [R]: columns solved
                for (int j = R - 1 - warp; j >= d; j -= kWarps) {
                    double t[72];
                    tile_fetch(tile_ptr(p, j, j - d), lane, t);
                    if (lane == 0) wait_at_least(flags + R, static_cast<unsigned int>(R - j));
                    __syncwarp();
                    xs[lane] = __ldcg(y + static_cast<size_t>(j) * NB + lane);
                    if (lane < NB - 32) xs[32 + lane] = __ldcg(y + static_cast<size_t>(j) * NB + 32 + lane);
                    __syncwarp();
                    double a, b, c;
                    tile_dot_regs(t, xs, lane, a, b, c);
                    const double a2 = __shfl_xor_sync(0xffffffffu, a, 16), b2 = __shfl_xor_sync(0xffffffffu, b, 16), c2 = __shfl_xor_sync(0xffffffffu, c, 16);
                    if (lane < 16) {
                        double* zc = zfar + static_cast<size_t>(j - d) * NB;
                        atomicAdd(zc + lane, a + b2);
                        atomicAdd(zc + lane + 16, c + a2);
                        atomicAdd(zc + lane + 32, b + c2);
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) red_release_add(flags + (j - d), 1u);
                }
            }
            return;
        }
        // ---------------------------------------------------------------------- CTA 0
        const int ring = kNear + 1;
        double* sD2 = dyn;                                        // third diagonal tile buffer (row-major [NB][NB])
        double* z = sD2 + NB * NB;                                // [ring][NB] pending sums of the near tiles
        double* sDinv = z + ring * NB;                            // [3][NB]
        double* sY = sDinv + 3 * NB;                              // [3][NB] y_j, travelling with the diagonal tile
        double* sW = sY + 3 * NB;                                 // [3][NB / 4][16] inverted diagonal blocks
        double* sZf = sW + 3 * (NB / 4) * 16;                     // [2][NB] the helpers' sums of a column
        uint64_t* bars = reinterpret_cast<uint64_t*>(sZf + 2 * NB);      // [3]: diagonal tile buffers
        double* sDiag[3] = {sA, sB, sD2};                         // sA / sB: row-major [NB][NB] here (the factor phases use stride kLd)
        uint32_t dphase = 0;
        long long bA = 0, bC = 0, bD = 0, bE = 0, bt = timing ? clock64() : 0;      // MSFM_BAND_DEBUG: phases of the loop below
#define BACK_TICK(acc) do { if (timing) { const long long t1 = clock64(); acc += t1 - bt; bt = t1; } } while (0)
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
            fence_mbar_init();
        }
        __syncthreads();
        for (int q = 0; q < p.nrhs; ++q) {
            double* y = p.y + static_cast<size_t>(q) * Npad;
            const double* zfar = p.zfar + static_cast<size_t>(q) * Npad;
            unsigned int* flags = p.flags + static_cast<size_t>(q) * (R + 1);
            // diagonal tile of column c with its pivots and y_c -> buffer c % 3 (one thread)
            auto fetch_diag = [&](int c) {
                const int b = c % 3;
                mbar_arrive_expect_tx(&bars[b], kTileBytes + 2 * NB * sizeof(double));
                bulk_g2s(sDiag[b], tile_ptr(p, c, c), kTileBytes, &bars[b]);
                bulk_g2s(sDinv + b * NB, p.dinv + static_cast<size_t>(c) * NB, NB * sizeof(double), &bars[b]);
                bulk_g2s(sY + b * NB, y + static_cast<size_t>(c) * NB, NB * sizeof(double), &bars[b]);
            };
            for (int i = threadIdx.x; i < ring * NB; i += kThreads) z[i] = 0.0;
            for (int i = threadIdx.x; i < 2 * NB; i += kThreads) sZf[i] = 0.0;        // the last column has no helper sums
            __syncthreads();
            if (warp == 1) {
                if (lane == 0) {
                    fence_proxy_async();
                    fetch_diag(R - 1);
                    if (R >= 2) fetch_diag(R - 2);
                }
                __syncwarp();
                const int b = (R - 1) % 3;
                mbar_wait(&bars[b], (dphase >> b) & 1u);          // the phase is consumed (toggled) when column R-1 is solved
                invert_diag_blocks(sDiag[b], sDinv + b * NB, sW + b * (NB / 4) * 16, lane);
            }
            __syncthreads();
            for (int j = R - 1; j >= 0; --j) {
                const int jb = j % 3;
                double* sx = sv + (j & 1) * NB;                   // x_j
                const double* sxp = sv + ((j + 1) & 1) * NB;      // x_{j+1}
                const int m_near = min(min(nbk, j), kNear);       // tiles (j, j - w), w = 1 .. m_near, take x_j here
                // (a) loads that do not depend on x_j
                double t[72];
                if (warp >= 1 && warp <= m_near) tile_fetch(tile_ptr(p, j, j - warp), lane, t);
                if (threadIdx.x == 7 * 32 && j >= 2) {
                    fence_proxy_async();
                    fetch_diag(j - 2);                             // two columns ahead: landed long before warp 1 inverts its blocks
                }
                // (c)
                if (warp == 0) {
                    const int l = lane;
                    double* zj = z + (j % ring) * NB;
                    const double* zf = sZf + (j & 1) * NB;
                    BACK_TICK(bE);
                    mbar_wait(&bars[jb], (dphase >> jb) & 1u);
                    BACK_TICK(bA);
                    const double* sYj = sY + jb * NB;
                    double b0 = sYj[l] - (zj[l] + zf[l]);
                    double b1 = l < NB - 32 ? sYj[32 + l] - (zj[32 + l] + zf[32 + l]) : 0.0;
                    zj[l] = 0.0;                                   // the slot now collects for column j - ring
                    if (l < NB - 32) zj[32 + l] = 0.0;
                    solve_diag_blocked(sDiag[jb], sW + jb * (NB / 4) * 16, l, b0, b1);
                    sx[l] = b0;
                    if (l < NB - 32) sx[32 + l] = b1;
                    BACK_TICK(bC);
                } else if (warp == 5 && j > 0) {
                    // the helpers' sums of column j - 1: complete when done[j-1] has counted every distance that reaches it
                    const int c = j - 1;
                    const int expect = max(min(nbk, R - 1 - c) - kNear, 0);
                    double* zf = sZf + (c & 1) * NB;
                    if (expect > 0) {
                        if (lane == 0) wait_at_least(flags + c, static_cast<unsigned int>(expect));
                        __syncwarp();
                        zf[lane] = __ldcg(zfar + static_cast<size_t>(c) * NB + lane);
                        if (lane < NB - 32) zf[32 + lane] = __ldcg(zfar + static_cast<size_t>(c) * NB + 32 + lane);
                    } else {
                        zf[lane] = 0.0;
                        if (lane < NB - 32) zf[32 + lane] = 0.0;
                    }
                } else if (warp == 6 && j + 1 < R) {
                    // publish x_{j+1} (solved a column ago) under this column's solve: the helpers and the caller read it from
                    // global memory; the fence and the release stay off the solver's path
                    double* yp = y + static_cast<size_t>(j + 1) * NB;
                    yp[lane] = sxp[lane];
                    if (lane < NB - 32) yp[32 + lane] = sxp[32 + lane];
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release(flags + R, static_cast<unsigned int>(R - 1 - j));
                }
                __syncthreads();
                // (d)
                if (warp >= 1 && warp <= m_near) {
                    double a, b, c;
                    tile_dot_regs(t, sx, lane, a, b, c);
                    tile_dot_commit(a, b, c, lane, z + ((j - warp) % ring) * NB);
                }
                if (warp == 7 && j > 0) {
                    // the next column's diagonal tile landed a column ago: invert its 4 x 4 blocks
                    const int b = (j - 1) % 3;
                    mbar_wait(&bars[b], (dphase >> b) & 1u);
                    invert_diag_blocks(sDiag[b], sDinv + b * NB, sW + b * (NB / 4) * 16, lane);
                }
                dphase ^= 1u << jb;                               // this column's diagonal buffer: phase consumed
                __syncthreads();                                  // z, sZf, sx and the diagonal buffer are free again
                if (warp == 0) BACK_TICK(bD);
            }
            if (warp == 6) {                                      // x_0
                y[lane] = sv[lane];
                if (lane < NB - 32) y[32 + lane] = sv[32 + lane];
            }
            __syncthreads();
        }
        if (timing) { p.dbg[5] = bA; p.dbg[6] = bC; p.dbg[7] = bD; p.dbg[8] = bE; }
#undef BACK_TICK
    }
    if (timing) {
        const long long t1 = clock64();
        p.dbg[0] = tD; p.dbg[1] = tP; p.dbg[2] = tU; p.dbg[3] = tS; p.dbg[4] = t1 - t0;
    }
#undef BAND_TICK
}

// block slot -> tile scatter of the fp32 blocks into the fp64 band (lower triangle of the renumbered matrix), damping added
__global__ void expand_band_kernel(ba::Problem P, const int32_t* __restrict__ pos, Params p, double inv_radius) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.n_blocks * 36) return;
    const int b = idx / 36, e = idx - 36 * b, i = e / 6, j = e - 6 * i;
    const int fa = __ldg(P.blk_row + b), fb = __ldg(P.blk_col + b);
    const int pa = __ldg(pos + fa), pb = __ldg(pos + fb);
    double v = static_cast<double>(P.sblk[idx]);                    // S[fa*6+i][fb*6+j]
    if (fa == fb) {
        if (j > i) return;                                         // lower triangle of a diagonal block, taken from its upper triangle
        v = static_cast<double>(P.sblk[b * 36 + 6 * j + i]);
        if (i == j) v += fmax(P.tail[P.tl.udiag + fa * 6 + i], 1e-6) * inv_radius;
    }
    int r, c;
    if (pa >= pb) { r = pa * 6 + i; c = pb * 6 + j; } else { r = pb * 6 + j; c = pa * 6 + i; }
    const int I = r / NB, J = c / NB;
    tile_ptr(p, I, J)[(r - I * NB) * NB + (c - J * NB)] = v;
}
// unit diagonal on the padding rows behind the last unknown
__global__ void pad_band_kernel(Params p, int n) {
    const int r = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R * NB) return;
    const int I = r / NB;
    tile_ptr(p, I, I)[(r - I * NB) * (NB + 1)] = 1.0;
}
__global__ void permute_in_kernel(const double* __restrict__ src, const int32_t* __restrict__ pos, int nf, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf * 6) return;
    const int f = i / 6, k = i - 6 * f;
    dst[static_cast<size_t>(pos[f]) * 6 + k] = src[i];
}
__global__ void permute_out_kernel(const double* __restrict__ src, const int32_t* __restrict__ pos, int nf, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf * 6) return;
    const int f = i / 6, k = i - 6 * f;
    dst[i] = src[static_cast<size_t>(pos[f]) * 6 + k];
}

struct BandSolver {
    int nf = 0, bw = 0, R = 0, nbk = 0, grid = 0;
    int32_t* d_pos = nullptr;
    double *tiles = nullptr, *y = nullptr, *dinv = nullptr;
    double* zfar = nullptr;         // [3][R * NB] doubles, then the back substitution's flags [3][R + 1] u32 (one allocation, one memset)
    size_t zfar_bytes = 0;
    int32_t* d_info = nullptr;      // [0] status, [2] the grid barrier counter
    long long* d_dbg = nullptr;
    size_t tile_bytes = 0;
};

}  // namespace band

namespace ba { void rcm_order(int, const std::vector<int32_t>&, const std::vector<int32_t>&, std::vector<int32_t>&, int&); }
using band::BandSolver;

// nullptr: the band is too wide for band storage to pay (the caller keeps the dense path), or an allocation failed (*err set)
BandSolver* band_create(int nf, const std::vector<int32_t>& blk_row, const std::vector<int32_t>& blk_col, int num_sms, cudaError_t* err) {
    *err = cudaSuccess;
    if (nf < 48) return nullptr;
    std::vector<int32_t> pos;
    int bw = 0;
    ba::rcm_order(nf, blk_row, blk_col, pos, bw);
    if (3 * std::max(1, bw) >= nf) return nullptr;
    BandSolver* B = new BandSolver();
    B->nf = nf; B->bw = bw;
    B->R = (6 * nf + band::NB - 1) / band::NB;
    B->nbk = 0;
    for (size_t k = 0; k < blk_row.size(); ++k) {
        const int pa = pos[blk_row[k]], pb = pos[blk_col[k]];
        const int hi = std::max(pa, pb) * 6 + 5, lo = std::min(pa, pb) * 6;
        B->nbk = std::max(B->nbk, hi / band::NB - lo / band::NB);
    }
    B->tile_bytes = static_cast<size_t>(B->R) * (B->nbk + 1) * band::NB * band::NB * sizeof(double);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&B->d_pos), static_cast<size_t>(nf) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&B->tiles), B->tile_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&B->y), 3 * static_cast<size_t>(B->R) * band::NB * sizeof(double));
    B->zfar_bytes = 3 * static_cast<size_t>(B->R) * band::NB * sizeof(double) + 3 * static_cast<size_t>(B->R + 1) * sizeof(unsigned int);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&B->zfar), B->zfar_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&B->dinv), static_cast<size_t>(B->R) * band::NB * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&B->d_info), 4 * sizeof(int32_t));
    if (e == cudaSuccess && getenv("MSFM_BAND_DEBUG")) e = cudaMalloc(reinterpret_cast<void**>(&B->d_dbg), 16 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpy(B->d_pos, pos.data(), static_cast<size_t>(nf) * sizeof(int32_t), cudaMemcpyHostToDevice);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(band::band_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(band::kDynSmem));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, band::band_cholesky_kernel, band::kThreads, band::kDynSmem);
    if (e == cudaSuccess && per_sm < 1) e = cudaErrorLaunchOutOfResources;
    if (e != cudaSuccess) {
        *err = e;
        cudaFree(B->d_pos); cudaFree(B->tiles); cudaFree(B->y); cudaFree(B->dinv); cudaFree(B->zfar); cudaFree(B->d_info);
        delete B;
        return nullptr;
    }
    // one CTA per SM, no more than the widest phase can use
    const int widest = B->nbk * (B->nbk + 1) / 2 + 1;
    B->grid = std::max(1, std::min(num_sms, widest));
    return B;
}
void band_destroy(BandSolver* B) {
    if (!B) return;
    if (B->d_dbg) {
        long long h[16] = {0};
        cudaMemcpy(h, B->d_dbg, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "band Cholesky, last solve, CTA 0 cycles: D %lld  P %lld  U %lld  barriers %lld  back substitution %lld  (%d tile columns)\n",
                h[0], h[1], h[2], h[3], h[4], B->R);
        fprintf(stderr, "  back substitution, thread 0: wait for the diagonal tile %lld  solve %lld  barrier + products + barrier %lld  loop head %lld\n", h[5], h[6], h[7], h[8]);
        cudaFree(B->d_dbg);
    }
    cudaFree(B->d_pos); cudaFree(B->tiles); cudaFree(B->y); cudaFree(B->dinv); cudaFree(B->zfar); cudaFree(B->d_info);
    delete B;
}
void band_info(const BandSolver* B, int32_t out[4]) { out[0] = B->bw; out[1] = band::NB; out[2] = B->R; out[3] = B->nbk; }
int32_t* band_dev_info(BandSolver* B) { return B->d_info; }

// Expand S (+ damping) into the band, factor, solve nrhs right-hand sides (columns of N = 6 nf doubles at stride N, caller's
// camera order; overwritten by the solutions).  Asynchronous on st; *band_dev_info != 0 afterwards = not positive definite.
cudaError_t band_factor_solve(BandSolver* B, const ba::Problem& P, double inv_radius, double* rhs, int nrhs, cudaStream_t st) {
    band::Params p;
    p.tiles = B->tiles; p.y = B->y; p.dinv = B->dinv; p.info = B->d_info; p.bar = reinterpret_cast<unsigned int*>(B->d_info + 2); p.dbg = B->d_dbg; p.zfar = B->zfar; p.flags = reinterpret_cast<unsigned int*>(B->zfar + 3 * static_cast<size_t>(B->R) * band::NB); p.R = B->R; p.nbk = B->nbk; p.nrhs = nrhs;
    const int N = 6 * B->nf, Npad = B->R * band::NB;
    cudaError_t e = cudaMemsetAsync(B->tiles, 0, B->tile_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(B->y, 0, 3 * static_cast<size_t>(Npad) * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(B->zfar, 0, B->zfar_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(B->d_info, 0, 4 * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    const int tot = P.n_blocks * 36;
    { band::expand_band_kernel<<<(tot + 255) / 256, 256, 0, st>>>(P, B->d_pos, p, inv_radius); MSFM_COUNT_LAUNCH(); }
    if (Npad > N) { band::pad_band_kernel<<<(Npad - N + 63) / 64, 64, 0, st>>>(p, N); MSFM_COUNT_LAUNCH(); }
    for (int c = 0; c < nrhs; ++c)
        { band::permute_in_kernel<<<(N + 255) / 256, 256, 0, st>>>(rhs + static_cast<size_t>(c) * N, B->d_pos, B->nf, B->y + static_cast<size_t>(c) * Npad); MSFM_COUNT_LAUNCH(); }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    void* args[] = {&p};
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(band::band_cholesky_kernel), dim3(B->grid), dim3(band::kThreads), args, band::kDynSmem, st);
    if (e != cudaSuccess) return e;
    MSFM_COUNT_LAUNCH();
    for (int c = 0; c < nrhs; ++c)
        { band::permute_out_kernel<<<(N + 255) / 256, 256, 0, st>>>(B->y + static_cast<size_t>(c) * Npad, B->d_pos, B->nf, rhs + static_cast<size_t>(c) * N); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

}  // namespace msfm
