// Batched geometric verification on the device: F-matrix RANSAC over the matches of many image pairs at once.
//
// Replaces the per-pair host call FeatureUtils::FilterMatches -> cv::findFundamentalMat(FM_RANSAC, 3.0, 0.99, mask)
// (src/Feature/FeatureUtils.cpp:176-206, called from FeatureMatcher::MatchImagePairs at src/Feature/FeatureMatching.cpp:60),
// which becomes the wall clock of ComputeMatches once the descriptor matching runs on the GPU.
// One CTA per image pair; the pair's aligned point pairs sit in shared memory.  The estimator is the one the host statement
// (host/src/GeometricVerification.cpp) defines: RANSAC over minimal 8-point samples with Hartley normalisation and the
// rank-2 constraint, OpenCV's error measure (the larger of the two squared point-to-epipolar-line distances against
// threshold^2), OpenCV's adaptive iteration count log(1 - confidence) / log(1 - w^8) capped at max_iters, then refits on the
// consensus set while they explain more points.  Hypotheses are generated from a counter-based generator, so the result is
// a function of the input only and independent of how the hypotheses are spread over threads: batches of 64 hypotheses are
// solved by 64 threads, scored by 8 warps (lanes over points), and folded into the running best IN HYPOTHESIS ORDER — exactly
// what a sequential loop over the same hypotheses would keep.  fp64 throughout.
// It cannot be bit-compatible with OpenCV's RANSAC (different sampler, 8- instead of 7-point minimal solver); the parity
// criterion is the agreement of the inlier sets (tests/test_verify_gpu.py).
#include <cuda_runtime.h>
#include <cstdint>
#include "launch_count.hpp"

namespace msfm {
namespace verify {

constexpr int kThreads = 256;
constexpr int kBatch = 64;            // hypotheses per batch
constexpr int kSmemPts = 2560;        // matches of a pair held in shared memory (more: read from global)
constexpr int kWsStride = 73;         // per-thread 8x9 scratch (doubles), odd stride

struct KpDev { const float2* xy; int32_t n; int32_t pad; };

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi with static indices (registers); returns the eigenvector of the
// smallest eigenvalue
__device__ inline void smallest_eigvec3(double a00, double a01, double a02, double a11, double a12, double a22, double v[3]) {
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
#pragma unroll 1
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-30) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = A[p][q];
            if (fabs(apq) < 1e-300) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
        }
    }
    int lo = 0;
    if (A[1][1] < A[0][0]) lo = 1;
    if (A[2][2] < (lo == 0 ? A[0][0] : A[1][1])) lo = 2;
    v[0] = lo == 0 ? V[0][0] : (lo == 1 ? V[0][1] : V[0][2]);
    v[1] = lo == 0 ? V[1][0] : (lo == 1 ? V[1][1] : V[1][2]);
    v[2] = lo == 0 ? V[2][0] : (lo == 1 ? V[2][1] : V[2][2]);
}

// f (unit 9-vector of the NORMALISED problem) -> rank 2 -> F = T2^T Fn T1 -> unit Frobenius norm.  false: degenerate.
__device__ inline bool finish_F(const double f[9], double c1x, double c1y, double s1, double c2x, double c2y, double s2, double F[9]) {
    // rank 2: remove the component along the weakest right singular vector (eigenvector of F^T F)
    double v[3];
    smallest_eigvec3(f[0] * f[0] + f[3] * f[3] + f[6] * f[6], f[0] * f[1] + f[3] * f[4] + f[6] * f[7], f[0] * f[2] + f[3] * f[5] + f[6] * f[8],
                     f[1] * f[1] + f[4] * f[4] + f[7] * f[7], f[1] * f[2] + f[4] * f[5] + f[7] * f[8], f[2] * f[2] + f[5] * f[5] + f[8] * f[8], v);
    double fr[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double d = f[3 * a] * v[0] + f[3 * a + 1] * v[1] + f[3 * a + 2] * v[2];
        fr[3 * a] = f[3 * a] - d * v[0]; fr[3 * a + 1] = f[3 * a + 1] - d * v[1]; fr[3 * a + 2] = f[3 * a + 2] - d * v[2];
    }
    // tmp = Fr T1,  T = [s 0 -s cx; 0 s -s cy; 0 0 1]
    double tmp[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        tmp[3 * a] = fr[3 * a] * s1;
        tmp[3 * a + 1] = fr[3 * a + 1] * s1;
        tmp[3 * a + 2] = -fr[3 * a] * s1 * c1x - fr[3 * a + 1] * s1 * c1y + fr[3 * a + 2];
    }
    // F = T2^T tmp
    double nrm = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        F[b] = s2 * tmp[b];
        F[3 + b] = s2 * tmp[3 + b];
        F[6 + b] = -s2 * c2x * tmp[b] - s2 * c2y * tmp[3 + b] + tmp[6 + b];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) nrm += F[i] * F[i];
    if (!(nrm > 1e-300) || !isfinite(nrm)) return false;
    nrm = 1.0 / sqrt(nrm);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] *= nrm;
    return true;
}

// OpenCV's error of a correspondence under F: the larger of the squared distances of x2 to the line F x1 and of x1 to
// the line F^T x2
__device__ inline double epipolar_error(const double F[9], float2 a, float2 b) {
    const double ax = a.x, ay = a.y, bx = b.x, by = b.y;
    const double l0 = F[0] * ax + F[1] * ay + F[2], l1 = F[3] * ax + F[4] * ay + F[5], l2 = F[6] * ax + F[7] * ay + F[8];
    const double d2 = bx * l0 + by * l1 + l2;
    const double e2 = d2 * d2 / (l0 * l0 + l1 * l1);
    const double m0 = F[0] * bx + F[3] * by + F[6], m1 = F[1] * bx + F[4] * by + F[7], m2 = F[2] * bx + F[5] * by + F[8];
    const double d1 = ax * m0 + ay * m1 + m2;
    const double e1 = d1 * d1 / (m0 * m0 + m1 * m1);
    return fmax(e1, e2);
}

struct PairPts {
    const float2* s1;       // shared copies (first n_s points) ...
    const float2* s2;
    int n_s;
    const float2* k1;       // ... the rest through the match list
    const float2* k2;
    const int2* m;
    __device__ inline void get(int i, float2& a, float2& b) const {
        if (i < n_s) { a = s1[i]; b = s2[i]; }
        else { const int2 q = m[i]; a = k1[q.x]; b = k2[q.y]; }
    }
};

__device__ inline double block_sum(double v, double* red /*[8]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    return s;
}

// Minimal solver: 8 sampled correspondences -> F.  ws: this thread's 8 x 9 scratch in shared memory.
__device__ inline bool eight_point_minimal(const PairPts& P, const int idx[8], double* ws, double F[9]) {
    float2 a[8], b[8];
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { P.get(idx[k], a[k], b[k]); c1x += a[k].x; c1y += a[k].y; c2x += b[k].x; c2y += b[k].y; }
    c1x *= 0.125; c1y *= 0.125; c2x *= 0.125; c2y *= 0.125;
    double d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        d1 += sqrt((a[k].x - c1x) * (a[k].x - c1x) + (a[k].y - c1y) * (a[k].y - c1y));
        d2 += sqrt((b[k].x - c2x) * (b[k].x - c2x) + (b[k].y - c2y) * (b[k].y - c2y));
    }
    d1 *= 0.125; d2 *= 0.125;
    const double s1 = d1 > 1e-12 ? 1.4142135623730951 / d1 : 1.0, s2 = d2 > 1e-12 ? 1.4142135623730951 / d2 : 1.0;
    double amax = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double x1 = (a[k].x - c1x) * s1, y1 = (a[k].y - c1y) * s1, x2 = (b[k].x - c2x) * s2, y2 = (b[k].y - c2y) * s2;
        double* r = ws + 9 * k;
        r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1.0;
    }
    // Gauss-Jordan with complete pivoting on the 8 x 9 system: the null vector has the one non-pivot column as free variable
    int colperm[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) colperm[j] = j;
    for (int r = 0; r < 8; ++r) {
        double best = -1.0;
        int pi = r, pj = r;
        for (int i = r; i < 8; ++i)
            for (int j = r; j < 9; ++j) {
                const double v = fabs(ws[9 * i + j]);
                if (v > best) { best = v; pi = i; pj = j; }
            }
        if (r == 0) amax = best;
        if (!(best > 1e-12 * amax)) return false;              // rank < 8: degenerate sample
        if (pi != r)
            for (int j = 0; j < 9; ++j) { const double t = ws[9 * r + j]; ws[9 * r + j] = ws[9 * pi + j]; ws[9 * pi + j] = t; }
        if (pj != r) {
            for (int i = 0; i < 8; ++i) { const double t = ws[9 * i + r]; ws[9 * i + r] = ws[9 * i + pj]; ws[9 * i + pj] = t; }
            const int t = colperm[r]; colperm[r] = colperm[pj]; colperm[pj] = t;
        }
        const double inv = 1.0 / ws[9 * r + r];
        for (int i = 0; i < 8; ++i) {
            if (i == r) continue;
            const double fct = ws[9 * i + r] * inv;
            for (int j = r + 1; j < 9; ++j) ws[9 * i + j] -= fct * ws[9 * r + j];
            ws[9 * i + r] = 0.0;
        }
    }
    double f[9];
    double nrm = 1.0;
    for (int j = 0; j < 9; ++j) f[j] = 0.0;
    for (int i = 0; i < 8; ++i) {
        const double x = -ws[9 * i + 8] / ws[9 * i + i];
        nrm += x * x;
        for (int j = 0; j < 9; ++j)
            if (colperm[i] == j) f[j] = x;
    }
    for (int j = 0; j < 9; ++j)
        if (colperm[8] == j) f[j] = 1.0;
    nrm = 1.0 / sqrt(nrm);
#pragma unroll
    for (int j = 0; j < 9; ++j) f[j] *= nrm;
    return finish_F(f, c1x, c1y, s1, c2x, c2y, s2, F);
}

// Symmetric 9 x 9 eigen-decomposition by cyclic Jacobi, executed by ONE WARP on shared memory (lane k < 9 owns index k of
// every rotation); eigenvalues on the diagonal of a, eigenvectors in the columns of v.
__device__ inline void jacobi9_warp(double* a, double* v, int lane) {
    for (int i = lane; i < 81; i += 32) v[i] = (i / 9 == i % 9) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 9; ++i)
            for (int j = i + 1; j < 9; ++j) off += a[9 * i + j] * a[9 * i + j];
        if (off < 1e-30) break;
        for (int p = 0; p < 9; ++p)
            for (int q = p + 1; q < 9; ++q) {
                const double apq = a[9 * p + q];
                if (fabs(apq) < 1e-300) continue;                 // warp-uniform
                const double theta = (a[9 * q + q] - a[9 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();
                if (lane < 9) {
                    const double akp = a[9 * lane + p], akq = a[9 * lane + q];
                    a[9 * lane + p] = c * akp - s * akq;
                    a[9 * lane + q] = s * akp + c * akq;
                }
                __syncwarp();
                if (lane < 9) {
                    const double apk = a[9 * p + lane], aqk = a[9 * q + lane];
                    a[9 * p + lane] = c * apk - s * aqk;
                    a[9 * q + lane] = s * apk + c * aqk;
                    const double vkp = v[9 * lane + p], vkq = v[9 * lane + q];
                    v[9 * lane + p] = c * vkp - s * vkq;
                    v[9 * lane + q] = s * vkp + c * vkq;
                }
                __syncwarp();
            }
    }
}

struct Shared {
    double F[kBatch][9];
    int cnt[kBatch];          // -1: degenerate sample
    double red[8];
    double ata[81], vec[81];
    double Fbest[9], Ftry[9];
    double norm[6];           // c1x c1y s1 c2x c2y s2
    int best, max_iters, n_in, flag;
};

__global__ void __launch_bounds__(kThreads)
ransac_pairs_kernel(const KpDev* __restrict__ kps, const int32_t* __restrict__ pair_slots /*[P][2]*/,
                    const long long* __restrict__ offsets /*[P+1]*/, const int32_t* __restrict__ matches /*[total][2]*/,
                    int P, double threshold, double confidence, int max_iters_opt,
                    uint8_t* __restrict__ mask /*[total]*/, int32_t* __restrict__ counts /*[P]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    float2* sp1 = reinterpret_cast<float2*>(smem_raw + ((sizeof(Shared) + 15) / 16) * 16);
    float2* sp2 = sp1 + kSmemPts;
    double* ws_all = reinterpret_cast<double*>(sp2 + kSmemPts);          // [kBatch][kWsStride]
    uint8_t* smask = reinterpret_cast<uint8_t*>(ws_all + kBatch * kWsStride);   // [kSmemPts] inlier flags of the current consensus set
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double thr2 = threshold * threshold;

    for (int pr = blockIdx.x; pr < P; pr += gridDim.x) {
        const long long beg = offsets[pr];
        const int n = static_cast<int>(offsets[pr + 1] - beg);
        uint8_t* gmask = mask + beg;
        const KpDev K1 = kps[pair_slots[2 * pr]], K2 = kps[pair_slots[2 * pr + 1]];
        PairPts pts;
        pts.m = reinterpret_cast<const int2*>(matches) + beg;
        pts.k1 = K1.xy; pts.k2 = K2.xy;
        pts.n_s = n < kSmemPts ? n : kSmemPts;
        pts.s1 = sp1; pts.s2 = sp2;
        __syncthreads();
        for (int i = threadIdx.x; i < pts.n_s; i += kThreads) {
            const int2 q = pts.m[i];
            sp1[i] = K1.xy[q.x];
            sp2[i] = K2.xy[q.y];
        }
        for (int i = threadIdx.x; i < n; i += kThreads) gmask[i] = 0;
        if (threadIdx.x == 0) { S.best = 0; S.max_iters = max_iters_opt; }
        __syncthreads();
        if (n < 8) {                        // cv::findFundamentalMat needs 8 points (7 for the minimal solver): nothing is kept
            if (threadIdx.x == 0 && counts) counts[pr] = 0;
            continue;
        }
        // ---- RANSAC over minimal samples
        for (int it0 = 0; it0 < S.max_iters; it0 += kBatch) {
            if (threadIdx.x < kBatch) {
                const int it = it0 + threadIdx.x;
                int idx[8];
                uint64_t ctr = static_cast<uint64_t>(it) << 8;
                for (int k = 0; k < 8;) {                            // 8 distinct indices
                    const int c = static_cast<int>(splitmix64(ctr++) % static_cast<uint64_t>(n));
                    bool dup = false;
                    for (int j = 0; j < k; ++j) dup = dup || idx[j] == c;
                    if (!dup) idx[k++] = c;
                }
                double F[9];
                const bool ok = eight_point_minimal(pts, idx, ws_all + threadIdx.x * kWsStride, F);
#pragma unroll
                for (int i = 0; i < 9; ++i) S.F[threadIdx.x][i] = F[i];
                S.cnt[threadIdx.x] = ok ? 0 : -1;
            }
            __syncthreads();
            // score: warp w takes hypotheses w, w + 8, ...; lanes over the points
            for (int h = warp; h < kBatch; h += kThreads / 32) {
                if (S.cnt[h] < 0 || it0 + h >= S.max_iters) continue;
                double F[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) F[i] = S.F[h][i];
                int c = 0;
                for (int i = lane; i < n; i += 32) {
                    float2 a, b;
                    pts.get(i, a, b);
                    c += epipolar_error(F, a, b) <= thr2 ? 1 : 0;      // NaN (degenerate line) compares false
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                __syncwarp();                                          // every lane has read S.cnt[h] (loop head) before lane 0 overwrites it
                if (lane == 0) S.cnt[h] = c;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                // fold the batch into the running best in hypothesis order (what a sequential loop would do)
                for (int h = 0; h < kBatch && it0 + h < S.max_iters; ++h) {
                    const int cnt = S.cnt[h];
                    if (cnt > S.best) {
                        S.best = cnt;
                        for (int i = 0; i < 9; ++i) S.Fbest[i] = S.F[h][i];
                        const int it = it0 + h;
                        // adaptive stop: enough samples to have drawn an all-inlier one with the requested confidence
                        const double w8 = pow(static_cast<double>(cnt) / n, 8.0);
                        if (w8 >= 1.0) {
                            S.max_iters = it + 1;
                        } else if (w8 > 1e-12) {
                            const double need = log(1.0 - confidence) / log1p(-w8);
                            if (need < static_cast<double>(S.max_iters)) {
                                const int nd = static_cast<int>(ceil(need));
                                S.max_iters = nd > it + 1 ? nd : it + 1;
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (S.best < 8) {
            if (threadIdx.x == 0 && counts) counts[pr] = 0;
            continue;
        }
        // ---- consensus set of the best model, then refits while they explain more points
        {
            double F[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) F[i] = S.Fbest[i];
            for (int i = threadIdx.x; i < n; i += kThreads) {
                float2 a, b;
                pts.get(i, a, b);
                gmask[i] = epipolar_error(F, a, b) <= thr2 ? 1 : 0;
            }
        }
        __syncthreads();
        for (int round = 0; round < 4; ++round) {
            // Hartley normalisation over the consensus set
            double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, cnt = 0;
            for (int i = threadIdx.x; i < n; i += kThreads)
                if (gmask[i]) { float2 a, b; pts.get(i, a, b); sx1 += a.x; sy1 += a.y; sx2 += b.x; sy2 += b.y; cnt += 1.0; }
            sx1 = block_sum(sx1, S.red); sy1 = block_sum(sy1, S.red); sx2 = block_sum(sx2, S.red); sy2 = block_sum(sy2, S.red);
            cnt = block_sum(cnt, S.red);
            if (cnt < 8.0) break;                                       // block-uniform
            const double c1x = sx1 / cnt, c1y = sy1 / cnt, c2x = sx2 / cnt, c2y = sy2 / cnt;
            double d1 = 0, d2 = 0;
            for (int i = threadIdx.x; i < n; i += kThreads)
                if (gmask[i]) {
                    float2 a, b; pts.get(i, a, b);
                    d1 += sqrt((a.x - c1x) * (a.x - c1x) + (a.y - c1y) * (a.y - c1y));
                    d2 += sqrt((b.x - c2x) * (b.x - c2x) + (b.y - c2y) * (b.y - c2y));
                }
            d1 = block_sum(d1, S.red) / cnt; d2 = block_sum(d2, S.red) / cnt;
            const double s1 = d1 > 1e-12 ? 1.4142135623730951 / d1 : 1.0, s2 = d2 > 1e-12 ? 1.4142135623730951 / d2 : 1.0;
            // A^T A (upper triangle, 45 sums)
            double acc[45];
#pragma unroll
            for (int k = 0; k < 45; ++k) acc[k] = 0.0;
            for (int i = threadIdx.x; i < n; i += kThreads)
                if (gmask[i]) {
                    float2 a, b; pts.get(i, a, b);
                    const double x1 = (a.x - c1x) * s1, y1 = (a.y - c1y) * s1, x2 = (b.x - c2x) * s2, y2 = (b.y - c2y) * s2;
                    const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
                    int k = 0;
#pragma unroll
                    for (int p = 0; p < 9; ++p)
#pragma unroll
                        for (int q = p; q < 9; ++q) acc[k++] += r[p] * r[q];
                }
            {
                int k = 0;
#pragma unroll
                for (int p = 0; p < 9; ++p)
#pragma unroll
                    for (int q = p; q < 9; ++q) {
                        const double v = block_sum(acc[k++], S.red);
                        if (threadIdx.x == 0) { S.ata[9 * p + q] = v; S.ata[9 * q + p] = v; }
                    }
            }
            __syncthreads();
            if (warp == 0) {
                jacobi9_warp(S.ata, S.vec, lane);
                if (lane == 0) {
                    int k = 0;
                    for (int i = 1; i < 9; ++i)
                        if (S.ata[10 * i] < S.ata[10 * k]) k = i;
                    double f[9], F[9];
                    for (int i = 0; i < 9; ++i) f[i] = S.vec[9 * i + k];
                    const bool ok = finish_F(f, c1x, c1y, s1, c2x, c2y, s2, F);
                    for (int i = 0; i < 9; ++i) S.Ftry[i] = F[i];
                    S.flag = ok ? 1 : 0;
                }
            }
            __syncthreads();
            if (!S.flag) break;
            double F[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) F[i] = S.Ftry[i];
            // count first (the refit is kept only if it explains MORE points), flags staged in shared / recomputed
            double c = 0.0;
            for (int i = threadIdx.x; i < n; i += kThreads) {
                float2 a, b; pts.get(i, a, b);
                const bool in = epipolar_error(F, a, b) <= thr2;
                if (i < kSmemPts) smask[i] = in ? 1 : 0;
                c += in ? 1.0 : 0.0;
            }
            c = block_sum(c, S.red);
            if (static_cast<int>(c) <= S.best) break;
            __syncthreads();
            if (threadIdx.x == 0) S.best = static_cast<int>(c);
            for (int i = threadIdx.x; i < n; i += kThreads) {
                if (i < kSmemPts) gmask[i] = smask[i];
                else { float2 a, b; pts.get(i, a, b); gmask[i] = epipolar_error(F, a, b) <= thr2 ? 1 : 0; }
            }
            __syncthreads();
        }
        __syncthreads();
        if (threadIdx.x == 0 && counts) counts[pr] = S.best;
    }
}

size_t ransac_smem_bytes() {
    return ((sizeof(Shared) + 15) / 16) * 16 + 2 * kSmemPts * sizeof(float2) + kBatch * kWsStride * sizeof(double) + kSmemPts;
}

}  // namespace verify

cudaError_t launch_ransac_pairs(const void* kps, const int32_t* pair_slots, const long long* offsets, const int32_t* matches, int P,
                                double threshold, double confidence, int max_iters, uint8_t* mask, int32_t* counts, int num_sms,
                                cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    const size_t smem = verify::ransac_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(verify::ransac_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int grid = P < num_sms * 2 ? P : num_sms * 2;
    { verify::ransac_pairs_kernel<<<grid, verify::kThreads, smem, st>>>(static_cast<const verify::KpDev*>(kps), pair_slots, offsets, matches, P,
                                                                      threshold, confidence, max_iters, mask, counts); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

}  // namespace msfm
