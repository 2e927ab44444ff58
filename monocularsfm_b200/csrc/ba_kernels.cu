// K2 — bundle-adjustment inner loop on the device.
//
// Replaces what Ceres does behind CeresBundelOptimizer::Optimize (src/Optimizer/CeresBundleOptimizer.cpp:293):
//   * per observation: the residual of BundleAutoDiffConstantFocalCostFunction (:29-53) and its 2 x (3+3+3)
//     Jacobian — here ANALYTIC, equal to the autodiff of Ceres' AngleAxisRotatePoint in both of its branches;
//   * per point: V = sum Jp^T Jp (+ Marquardt damping), g_p; per camera: U, g_c;
//   * Schur elimination of the points onto the free cameras: S = U + D_c - sum W V^-1 W^T, rhs = -(g_c - W V^-1 g_p)
//     (Ceres SchurEliminator for DENSE_SCHUR / SPARSE_SCHUR, :264-273) into one dense buffer
//     [S | rhs | g_c | diag U | cost] that a single NCCL all-reduce sums across GPUs.  The reduction is a GATHER, not
//     a scatter: the sparsity pattern is fixed across LM iterations, so per-camera observation lists and per
//     camera-pair co-observation lists are built once per problem; each diagonal block is then owned by one CTA and
//     each off-diagonal block by one warp, which sum their contributions in registers and store — no floating-point
//     atomics (the first version scattered 1.1e8 fp64 atomics per linearisation and was bound by L2 atomic
//     throughput: profiles/r01_k2_ncu_raw.csv);
//   * back-substitution of the points, candidate-step evaluation.
// Observations are grouped by point; one warp owns one point, one lane one observation (chunks of 32 for longer
// tracks).  Geometry (projection, residual, V^-1) is evaluated in fp64, the 6x3 / 6x6 block products of the Schur
// reduction in fp32, all accumulation into the normal equations in fp64.  Tensor cores are not used: the work is
// a sparse gather/scatter bounded by HBM/L2 atomics, not by flops.
#include <cuda_runtime.h>
#include <cstdint>

#include "ba_types.cuh"

namespace msfm {
namespace ba {

// ------------------------------------------------------------------------------------------------ per camera
// a = sin(th)/th, b = (1-cos th)/th^2, a1 = (th cos th - sin th)/th^3, b1 = (th sin th - 2(1-cos th))/th^4
// with series below th^2 = 0.25 (no cancellation) and Ceres' Taylor branch (th^2 <= DBL_EPSILON): R x = x + w x x.
__global__ void cam_prep_kernel(const double* __restrict__ cams, int n_cams, CamPre* __restrict__ pre) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cams) return;
    const double wx = cams[6 * c + 0], wy = cams[6 * c + 1], wz = cams[6 * c + 2];
    const double th2 = wx * wx + wy * wy + wz * wz;
    double a, b, a1, b1;
    if (!(th2 > 2.220446049250313e-16)) {
        a = 1.0; b = 0.0; a1 = 0.0; b1 = 0.0;                      // Ceres' first-order branch
    } else if (th2 < 0.25) {
        const double t2 = th2, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4, t10 = t8 * t2;
        a = 1.0 - t2 / 6.0 + t4 / 120.0 - t6 / 5040.0 + t8 / 362880.0 - t10 / 39916800.0;
        b = 0.5 - t2 / 24.0 + t4 / 720.0 - t6 / 40320.0 + t8 / 3628800.0 - t10 / 479001600.0;
        a1 = -1.0 / 3.0 + t2 / 30.0 - t4 / 840.0 + t6 / 45360.0 - t8 / 3991680.0 + t10 / 518918400.0;
        b1 = -1.0 / 12.0 + t2 / 180.0 - t4 / 6720.0 + t6 / 453600.0 - t8 / 47900160.0 + t10 / 7264857600.0;
    } else {
        const double th = sqrt(th2), s = sin(th), co = cos(th);
        a = s / th;
        b = (1.0 - co) / th2;
        a1 = (th * co - s) / (th2 * th);
        b1 = (th * s - 2.0 * (1.0 - co)) / (th2 * th2);
    }
    CamPre p;
    p.w[0] = wx; p.w[1] = wy; p.w[2] = wz;
    p.t[0] = cams[6 * c + 3]; p.t[1] = cams[6 * c + 4]; p.t[2] = cams[6 * c + 5];
    p.a = a; p.b = b; p.a1 = a1; p.b1 = b1;
    // R = I + a [w]x + b [w]x^2 ,  [w]x^2 = w w^T - th2 I
    p.R[0] = 1.0 + b * (wx * wx - th2); p.R[1] = -a * wz + b * wx * wy;      p.R[2] = a * wy + b * wx * wz;
    p.R[3] = a * wz + b * wx * wy;      p.R[4] = 1.0 + b * (wy * wy - th2); p.R[5] = -a * wx + b * wy * wz;
    p.R[6] = -a * wy + b * wx * wz;     p.R[7] = a * wx + b * wy * wz;      p.R[8] = 1.0 + b * (wz * wz - th2);
    pre[c] = p;
}

// ------------------------------------------------------------------------------------------------ per observation
struct ObsLin {
    double r[2];
    float Jc[12];   // [2][6]  d r / d (rvec | tvec)
    float Jp[6];    // [2][3]  d r / d point
};

template <bool kJac>
__device__ __forceinline__ void obs_eval(const CamPre& c, const double X[3], double u, double v, double fx, double fy,
                                         double r[2], double Jc[12], double Jp[6], double xy[2] = nullptr) {
    const double px = c.R[0] * X[0] + c.R[1] * X[1] + c.R[2] * X[2] + c.t[0];
    const double py = c.R[3] * X[0] + c.R[4] * X[1] + c.R[5] * X[2] + c.t[1];
    const double pz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
    const double iz = 1.0 / pz;
    const double xp = px * iz, yp = py * iz;
    r[0] = fx * xp - u;
    r[1] = fy * yp - v;
    if (xy) { xy[0] = xp; xy[1] = yp; }       // d r / d (fx, fy) = diag(xp, yp)
    if (!kJac) return;
    // A = d(u,v)/dp
    const double A00 = fx * iz, A02 = -fx * xp * iz, A11 = fy * iz, A12 = -fy * yp * iz;
    // d(R x)/dw = -a [X]x + a1 c w^T - b ([c]x + X w^T - (w.X) I) + b1 e w^T,  c = w x X, e = w x c
    const double wx = c.w[0], wy = c.w[1], wz = c.w[2];
    const double cx = wy * X[2] - wz * X[1], cy = wz * X[0] - wx * X[2], cz = wx * X[1] - wy * X[0];
    const double ex = wy * cz - wz * cy, ey = wz * cx - wx * cz, ez = wx * cy - wy * cx;
    const double wX = wx * X[0] + wy * X[1] + wz * X[2];
    const double a = c.a, b = c.b, a1 = c.a1, b1 = c.b1;
    double M[9];
    const double vx = a1 * cx + b1 * ex - b * X[0], vy = a1 * cy + b1 * ey - b * X[1], vz = a1 * cz + b1 * ez - b * X[2];
    // rank-one part (a1 c + b1 e - b X) w^T plus b (w.X) I
    M[0] = vx * wx + b * wX; M[1] = vx * wy;          M[2] = vx * wz;
    M[3] = vy * wx;          M[4] = vy * wy + b * wX; M[5] = vy * wz;
    M[6] = vz * wx;          M[7] = vz * wy;          M[8] = vz * wz + b * wX;
    // -a [X]x - b [c]x  with [q]x = [0 -qz qy; qz 0 -qx; -qy qx 0]
    const double qx = a * X[0] + b * cx, qy = a * X[1] + b * cy, qz = a * X[2] + b * cz;
    M[1] += qz;  M[2] -= qy;
    M[3] -= qz;  M[5] += qx;
    M[6] += qy;  M[7] -= qx;
    // Jc = [A M | A]
    Jc[0] = A00 * M[0] + A02 * M[6]; Jc[1] = A00 * M[1] + A02 * M[7]; Jc[2] = A00 * M[2] + A02 * M[8];
    Jc[3] = A00; Jc[4] = 0.0; Jc[5] = A02;
    Jc[6] = A11 * M[3] + A12 * M[6]; Jc[7] = A11 * M[4] + A12 * M[7]; Jc[8] = A11 * M[5] + A12 * M[8];
    Jc[9] = 0.0; Jc[10] = A11; Jc[11] = A12;
    // Jp = A R
    Jp[0] = A00 * c.R[0] + A02 * c.R[6]; Jp[1] = A00 * c.R[1] + A02 * c.R[7]; Jp[2] = A00 * c.R[2] + A02 * c.R[8];
    Jp[3] = A11 * c.R[3] + A12 * c.R[6]; Jp[4] = A11 * c.R[4] + A12 * c.R[7]; Jp[5] = A11 * c.R[5] + A12 * c.R[8];
}

// residual + Jacobian dump for parity tests (msfm_ba_evaluate)
__global__ void evaluate_kernel(const CamPre* __restrict__ pre, const double* __restrict__ pts,
                                const double* __restrict__ obs_uv, const int32_t* __restrict__ obs_cam,
                                const int32_t* __restrict__ obs_pt, int n_obs, double fx, double fy,
                                double* __restrict__ r_out, float* __restrict__ J_out, double* __restrict__ cost) {
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += gridDim.x * blockDim.x) {
        const CamPre c = pre[obs_cam[i]];
        const int p = obs_pt[i];
        const double X[3] = {pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]};
        double r[2], Jc[12], Jp[6];
        if (J_out) obs_eval<true>(c, X, obs_uv[2 * i], obs_uv[2 * i + 1], fx, fy, r, Jc, Jp);
        else obs_eval<false>(c, X, obs_uv[2 * i], obs_uv[2 * i + 1], fx, fy, r, Jc, Jp);
        local += 0.5 * (r[0] * r[0] + r[1] * r[1]);
        if (r_out) { r_out[2 * i] = r[0]; r_out[2 * i + 1] = r[1]; }
        if (J_out) {
            float* J = J_out + static_cast<size_t>(i) * 18;
#pragma unroll
            for (int k = 0; k < 6; ++k) { J[k] = static_cast<float>(Jc[k]); J[9 + k] = static_cast<float>(Jc[6 + k]); }
#pragma unroll
            for (int k = 0; k < 3; ++k) { J[6 + k] = static_cast<float>(Jp[k]); J[15 + k] = static_cast<float>(Jp[3 + k]); }
        }
    }
    // block reduce -> one fp64 atomic per block
    __shared__ double sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && cost) atomicAdd(cost, v);
    }
}

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double warp_sum(double v);

// Mean reprojection error of every point over its observations, sqrt(dx^2 + dy^2) per observation: what
// Map::UpdateFromBAData recomputes on the host after every BA through ComputeTrackError
// (src/Reconstruction/Map.cpp:1201, 1834-1846; Projection::CalculateReprojectionError, Projection.cpp:114-133).
// One warp per point, one lane per observation.
__global__ void __launch_bounds__(256)
track_error_kernel(Problem P, double* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < P.n_pts; p += gridDim.x * wpb) {
        const int beg = P.pt_start[p], end = P.pt_start[p + 1];
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        double s = 0.0;
        for (int o = beg + lane; o < end; o += 32) {
            const CamPre c = P.pre[P.obs_cam[o]];
            double r[2], Jc[12], Jp[6];
            obs_eval<false>(c, X, P.obs_uv[2 * o], P.obs_uv[2 * o + 1], P.fx, P.fy, r, Jc, Jp);
            s += sqrt(r[0] * r[0] + r[1] * r[1]);
        }
        s = warp_sum(s);
        if (lane == 0) err[p] = end > beg ? s / static_cast<double>(end - beg) : 0.0;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inverse of the symmetric 3x3 V (v00 v01 v02 v11 v12 v22)
__device__ __forceinline__ void sym3_inverse(const double v[6], double inv[6]) {
    const double c00 = v[3] * v[5] - v[4] * v[4];
    const double c01 = v[2] * v[4] - v[1] * v[5];
    const double c02 = v[1] * v[4] - v[2] * v[3];
    const double det = v[0] * c00 + v[1] * c01 + v[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
    inv[3] = (v[0] * v[5] - v[2] * v[2]) * id;
    inv[4] = (v[1] * v[2] - v[0] * v[4]) * id;
    inv[5] = (v[0] * v[3] - v[1] * v[1]) * id;
}

struct LaneObs {           // one observation in a lane
    bool valid;
    int cam, f;            // camera index, reduced (free) index or -1
    double r[2];
    double xy[2];          // (xp, yp): the focal Jacobian
    float Jc[12], Jp[6];
};

__device__ __forceinline__ void load_lane(const Problem& P, int obs, bool valid, const double X[3], LaneObs& o) {
    o.valid = valid;
    o.cam = -1; o.f = -1;
    o.r[0] = o.r[1] = 0.0;
    o.xy[0] = o.xy[1] = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) o.Jc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) o.Jp[k] = 0.f;
    if (!valid) return;
    o.cam = P.obs_cam[obs];
    o.f = P.cam_free[o.cam];
    const CamPre c = P.pre[o.cam];
    double Jc[12], Jp[6];
    obs_eval<true>(c, X, P.obs_uv[2 * obs], P.obs_uv[2 * obs + 1], P.fx, P.fy, o.r, Jc, Jp, o.xy);
    if (o.f >= 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) o.Jc[k] = static_cast<float>(Jc[k]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) o.Jp[k] = static_cast<float>(Jp[k]);
}

// per-point normal-equation pieces: V (damped), V^-1, g_p — reduced over the whole track
// kFocal: also wf = Wf = sum Jf^T Jp (2x3) and fstat = sum xp^2, sum yp^2, sum xp r0, sum yp r1 over the track.
// A template parameter, not a run-time flag: the extra accumulators cost the constant-focal kernel 70 registers
// (118 -> 188, one CTA per SM instead of two) when they were merely branched around.
template <bool kFocal = false>
__device__ __forceinline__ void point_pass1(const Problem& P, int p, int beg, int end, int lane, const double X[3],
                                            double inv_radius, double Vinv[6], double gp[3], LaneObs& first,
                                            double* wf = nullptr, double* fstat = nullptr) {
    double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    double w6[6] = {0, 0, 0, 0, 0, 0}, f4[4] = {0, 0, 0, 0};
    for (int base = beg; base < end; base += 32) {
        LaneObs o;
        load_lane(P, base + lane, base + lane < end, X, o);
        if (base == beg) first = o;
        const double j0 = o.Jp[0], j1 = o.Jp[1], j2 = o.Jp[2], j3 = o.Jp[3], j4 = o.Jp[4], j5 = o.Jp[5];
        v[0] += j0 * j0 + j3 * j3; v[1] += j0 * j1 + j3 * j4; v[2] += j0 * j2 + j3 * j5;
        v[3] += j1 * j1 + j4 * j4; v[4] += j1 * j2 + j4 * j5; v[5] += j2 * j2 + j5 * j5;
        g[0] += j0 * o.r[0] + j3 * o.r[1]; g[1] += j1 * o.r[0] + j4 * o.r[1]; g[2] += j2 * o.r[0] + j5 * o.r[1];
        if (kFocal) {
            w6[0] += o.xy[0] * j0; w6[1] += o.xy[0] * j1; w6[2] += o.xy[0] * j2;
            w6[3] += o.xy[1] * j3; w6[4] += o.xy[1] * j4; w6[5] += o.xy[1] * j5;
        }
        if (kFocal) {
            f4[0] += o.xy[0] * o.xy[0]; f4[1] += o.xy[1] * o.xy[1];
            f4[2] += o.xy[0] * o.r[0];  f4[3] += o.xy[1] * o.r[1];
        }
    }
    if (kFocal) {
#pragma unroll
        for (int k = 0; k < 6; ++k) wf[k] = warp_sum(w6[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) fstat[k] = warp_sum(f4[k]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) gp[k] = warp_sum(g[k]);
    // Marquardt damping D^2 = max(diag, 1e-6) / radius  (Ceres LM strategy with Jacobi scaling, min_lm_diagonal)
    v[0] += fmax(v[0], 1e-6) * inv_radius;
    v[3] += fmax(v[3], 1e-6) * inv_radius;
    v[5] += fmax(v[5], 1e-6) * inv_radius;
    sym3_inverse(v, Vinv);
}

// ------------------------------------------------------------------------------------------------ linearize + Schur
// sys layout (fp64): S [n6*n6] | rhs [n6] | gc [n6] | udiag [n6] | scalars[8] (0: cost)

// Pass P — one warp per point: residuals + Jacobians of its observations (stored), damped V^-1 and g_p (stored),
// cost and max |g_p|.
template <bool kFocal>
__global__ void __launch_bounds__(256)
point_pass_kernel(Problem P, double inv_radius, double* __restrict__ sys) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const size_t n6 = static_cast<size_t>(P.n_free) * 6;
    double* scal = sys + n6 * n6 + 3 * n6;
    double cost_local = 0.0, gpmax_local = 0.0;
    double ff[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};        // focal block sums of this warp (lane 0): F00 F01 F11 rhsf0 rhsf1 gf0 gf1 uf0 uf1
    constexpr bool focal = kFocal;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < P.n_pts; p += gridDim.x * wpb) {
        const int beg = P.pt_start[p], end = P.pt_start[p + 1];
        if (beg == end) {
            if (focal && lane < 6) P.pt_Wf[6 * static_cast<size_t>(p) + lane] = 0.0;
            continue;
        }
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        double Vinv[6], gp[3], Wf[6], fs[4];
        LaneObs A;
        point_pass1<kFocal>(P, p, beg, end, lane, X, inv_radius, Vinv, gp, A, Wf, fs);
        if (focal) {
            // T = Wf V^-1 (2x3);  F -= T Wf^T,  rhs_f += T g_p - g_f   (the point is eliminated from the focal block too)
            if (lane < 6) P.pt_Wf[6 * static_cast<size_t>(p) + lane] = Wf[lane];
            const double T0[3] = {Wf[0] * Vinv[0] + Wf[1] * Vinv[1] + Wf[2] * Vinv[2], Wf[0] * Vinv[1] + Wf[1] * Vinv[3] + Wf[2] * Vinv[4],
                                  Wf[0] * Vinv[2] + Wf[1] * Vinv[4] + Wf[2] * Vinv[5]};
            const double T1[3] = {Wf[3] * Vinv[0] + Wf[4] * Vinv[1] + Wf[5] * Vinv[2], Wf[3] * Vinv[1] + Wf[4] * Vinv[3] + Wf[5] * Vinv[4],
                                  Wf[3] * Vinv[2] + Wf[4] * Vinv[4] + Wf[5] * Vinv[5]};
            ff[0] += fs[0] - (T0[0] * Wf[0] + T0[1] * Wf[1] + T0[2] * Wf[2]);
            ff[1] += -(T0[0] * Wf[3] + T0[1] * Wf[4] + T0[2] * Wf[5]);
            ff[2] += fs[1] - (T1[0] * Wf[3] + T1[1] * Wf[4] + T1[2] * Wf[5]);
            ff[3] += (T0[0] * gp[0] + T0[1] * gp[1] + T0[2] * gp[2]) - fs[2];
            ff[4] += (T1[0] * gp[0] + T1[1] * gp[1] + T1[2] * gp[2]) - fs[3];
            ff[5] += fs[2]; ff[6] += fs[3]; ff[7] += fs[0]; ff[8] += fs[1];
        }
        gpmax_local = fmax(gpmax_local, fmax(fabs(gp[0]), fmax(fabs(gp[1]), fabs(gp[2]))));
        if (lane < 6) P.pt_Vinv[6 * static_cast<size_t>(p) + lane] = Vinv[lane];
        if (lane < 3) P.pt_gp[3 * static_cast<size_t>(p) + lane] = gp[lane];
        for (int base = beg; base < end; base += 32) {
            if (base != beg) load_lane(P, base + lane, base + lane < end, X, A);
            if (!A.valid) continue;
            cost_local += 0.5 * (A.r[0] * A.r[0] + A.r[1] * A.r[1]);
            const size_t o = static_cast<size_t>(base + lane);
            float2* J2 = reinterpret_cast<float2*>(P.obs_J + o * 18);          // 72-byte rows are 8-byte aligned
#pragma unroll
            for (int k = 0; k < 6; ++k) J2[k] = make_float2(A.Jc[2 * k], A.Jc[2 * k + 1]);
#pragma unroll
            for (int k = 0; k < 3; ++k) J2[6 + k] = make_float2(A.Jp[2 * k], A.Jp[2 * k + 1]);
            reinterpret_cast<double2*>(P.obs_r)[o] = make_double2(A.r[0], A.r[1]);
        }
    }
    __shared__ double sh_c[8], sh_g[8];
    cost_local = warp_sum(cost_local);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gpmax_local = fmax(gpmax_local, __shfl_xor_sync(0xffffffffu, gpmax_local, o));
    if (lane == 0) { sh_c[threadIdx.x >> 5] = cost_local; sh_g[threadIdx.x >> 5] = gpmax_local; }
    __syncthreads();
    if (focal) {       // 9 fp64 atomics per CTA into the focal slots behind the scalars
        __shared__ double sh_f[8][9];
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) sh_f[threadIdx.x >> 5][k] = ff[k];
        }
        __syncthreads();
        if (threadIdx.x < 9) {
            double v = 0.0;
            for (int k = 0; k < wpb; ++k) v += sh_f[k][threadIdx.x];
            atomicAdd(scal + 8 + 2 * n6 + threadIdx.x, v);
        }
    }
    if (threadIdx.x == 0) {
        double c = 0.0, g = 0.0;
        for (int k = 0; k < wpb; ++k) { c += sh_c[k]; g = fmax(g, sh_g[k]); }
        atomicAdd(&scal[0], c);          // one fp64 atomic per CTA (cost only)
        // max of non-negative doubles == max of their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(P.gpmax_bits), static_cast<unsigned long long>(__double_as_longlong(g)));
    }
}

__device__ __forceinline__ void load_obs_J(const float* __restrict__ obs_J, int o, float Jc[12], float Jp[6]) {
    const float2* J2 = reinterpret_cast<const float2*>(obs_J + static_cast<size_t>(o) * 18);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float2 v = __ldg(J2 + k); Jc[2 * k] = v.x; Jc[2 * k + 1] = v.y; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { const float2 v = __ldg(J2 + 6 + k); Jp[2 * k] = v.x; Jp[2 * k + 1] = v.y; }
}
// W = Jc^T Jp (6x3), Y = W V^-1 (6x3) with the symmetric V^-1 = (v0 v1 v2 / v1 v3 v4 / v2 v4 v5)
__device__ __forceinline__ void make_WY(const float Jc[12], const float Jp[6], const float vi[6], float W[18], float Y[18]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k) W[3 * i + k] = Jc[i] * Jp[k] + Jc[6 + i] * Jp[3 + k];
        Y[3 * i + 0] = W[3 * i] * vi[0] + W[3 * i + 1] * vi[1] + W[3 * i + 2] * vi[2];
        Y[3 * i + 1] = W[3 * i] * vi[1] + W[3 * i + 1] * vi[3] + W[3 * i + 2] * vi[4];
        Y[3 * i + 2] = W[3 * i] * vi[2] + W[3 * i + 1] * vi[4] + W[3 * i + 2] * vi[5];
    }
}

// Pass C — one CTA per free camera: diagonal block U_c - sum Y W^T, rhs_c, g_c, diag U_c over the camera's observations.
__global__ void __launch_bounds__(256)
camera_diag_kernel(Problem P, double* __restrict__ sys) {
    const int f = blockIdx.x;
    if (f >= P.n_free) return;
    const size_t n6 = static_cast<size_t>(P.n_free) * 6;
    double acc[54];                      // 36 block | 6 rhs | 6 gc | 6 udiag
#pragma unroll
    for (int k = 0; k < 54; ++k) acc[k] = 0.0;
    const int beg = P.cam_obs_start[f], end = P.cam_obs_start[f + 1];
    for (int idx = beg + threadIdx.x; idx < end; idx += blockDim.x) {
        const int o = __ldg(P.cam_obs_list + idx);
        const int p = __ldg(P.obs_pt + o);
        float Jc[12], Jp[6], vi[6], W[18], Y[18];
        load_obs_J(P.obs_J, o, Jc, Jp);
        double gp[3];
#pragma unroll
        for (int k = 0; k < 6; ++k) vi[k] = static_cast<float>(__ldg(P.pt_Vinv + 6 * static_cast<size_t>(p) + k));
#pragma unroll
        for (int k = 0; k < 3; ++k) gp[k] = __ldg(P.pt_gp + 3 * static_cast<size_t>(p) + k);
        const double2 r = __ldg(reinterpret_cast<const double2*>(P.obs_r) + o);
        make_WY(Jc, Jp, vi, W, Y);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double jr = static_cast<double>(Jc[i]) * r.x + static_cast<double>(Jc[6 + i]) * r.y;
            acc[36 + i] += Y[3 * i] * gp[0] + Y[3 * i + 1] * gp[1] + Y[3 * i + 2] * gp[2] - jr;
            acc[42 + i] += jr;
            acc[48 + i] += static_cast<double>(Jc[i] * Jc[i] + Jc[6 + i] * Jc[6 + i]);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float u = Jc[i] * Jc[j] + Jc[6 + i] * Jc[6 + j];
                const float yw = Y[3 * i] * W[3 * j] + Y[3 * i + 1] * W[3 * j + 1] + Y[3 * i + 2] * W[3 * j + 2];
                acc[6 * i + j] += static_cast<double>(u - yw);
            }
        }
    }
    __shared__ double sh[8][54];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 54; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 54) {
        double v = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) v += sh[w][threadIdx.x];
        const int k = threadIdx.x;
        double* S = sys;
        double* rhs = S + n6 * n6;
        if (k < 36) S[(static_cast<size_t>(f) * 6 + k / 6) * n6 + static_cast<size_t>(f) * 6 + k % 6] = v;
        else if (k < 42) rhs[f * 6 + (k - 36)] = v;
        else if (k < 48) rhs[n6 + f * 6 + (k - 42)] = v;           // gc
        else rhs[2 * n6 + f * 6 + (k - 48)] = v;                   // udiag
    }
}

// Pass F (refine_focal only) — one CTA per free camera: the 6 x 2 border block that couples the camera with the shared
// focal block, B_c = sum_obs (Jc^T Jf - Y Wf_p^T), Jf = diag(xp, yp) recovered from the stored Jacobian
// (Jc[3] = fx/pz, Jc[5] = -fx xp/pz;  Jc[10] = fy/pz, Jc[11] = -fy yp/pz).  Stored as two columns B0 | B1 behind the
// scalars of the system buffer.
__global__ void __launch_bounds__(256)
camera_focal_border_kernel(Problem P, double* __restrict__ sys) {
    const int f = blockIdx.x;
    if (f >= P.n_free) return;
    const size_t n6 = static_cast<size_t>(P.n_free) * 6;
    double acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0;
    const int beg = P.cam_obs_start[f], end = P.cam_obs_start[f + 1];
    for (int idx = beg + threadIdx.x; idx < end; idx += blockDim.x) {
        const int o = __ldg(P.cam_obs_list + idx);
        const int p = __ldg(P.obs_pt + o);
        float Jc[12], Jp[6], vi[6], W[18], Y[18], wf[6];
        load_obs_J(P.obs_J, o, Jc, Jp);
#pragma unroll
        for (int k = 0; k < 6; ++k) vi[k] = static_cast<float>(__ldg(P.pt_Vinv + 6 * static_cast<size_t>(p) + k));
#pragma unroll
        for (int k = 0; k < 6; ++k) wf[k] = static_cast<float>(__ldg(P.pt_Wf + 6 * static_cast<size_t>(p) + k));
        make_WY(Jc, Jp, vi, W, Y);
        const float xp = -Jc[5] / Jc[3], yp = -Jc[11] / Jc[10];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            acc[2 * i] += static_cast<double>(Jc[i] * xp - (Y[3 * i] * wf[0] + Y[3 * i + 1] * wf[1] + Y[3 * i + 2] * wf[2]));
            acc[2 * i + 1] += static_cast<double>(Jc[6 + i] * yp - (Y[3 * i] * wf[3] + Y[3 * i + 1] * wf[4] + Y[3 * i + 2] * wf[5]));
        }
    }
    __shared__ double sh[8][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double v = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) v += sh[w][threadIdx.x];
        double* B = sys + n6 * n6 + 3 * n6 + 8;
        B[(threadIdx.x & 1) * n6 + static_cast<size_t>(f) * 6 + (threadIdx.x >> 1)] = v;
    }
}

// Pass B — one CTA (4 warps) per non-empty camera-pair block (fa < fb): - sum over co-observing points of Y_a W_b^T.
// The tuples of a block are spread over the 128 threads (block-stride), summed in fp32 registers, reduced by warp
// shuffles and combined across the four warps in a fixed order: plain stores, deterministic.  The first version gave a
// whole block to ONE warp (grid = all n_free^2 blocks, most of them empty): ~2500 busy warps with ~1000 tuples each
// were bound by the latency of the dependent gathers (ncu: sm throughput 11 %, L2 22 %; 418 of the 540 us of a
// linearisation at configs[3], profiles/r01b_k2_ncu_summary.txt).
constexpr int kPairBlockThreads = 128;
__global__ void __launch_bounds__(kPairBlockThreads)
pair_block_kernel(Problem P, double* __restrict__ sys) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t n6 = static_cast<size_t>(P.n_free) * 6;
    __shared__ float part[kPairBlockThreads / 32][36];
    for (int bi = blockIdx.x; bi < P.n_blk_list; bi += gridDim.x) {
        const int b = __ldg(P.blk_list + bi);
        const int beg = __ldg(P.blk_start + b), end = __ldg(P.blk_start + b + 1);
        float acc[36];
#pragma unroll
        for (int k = 0; k < 36; ++k) acc[k] = 0.f;
        for (int idx = beg + threadIdx.x; idx < end; idx += kPairBlockThreads) {
            const int2 t = __ldg(P.blk_tuples + idx);
            const int p = __ldg(P.obs_pt + t.x);
            float Jca[12], Jpa[6], Jcb[12], Jpb[6], vi[6], W[18], Y[18];
            load_obs_J(P.obs_J, t.x, Jca, Jpa);
            load_obs_J(P.obs_J, t.y, Jcb, Jpb);
#pragma unroll
            for (int k = 0; k < 6; ++k) vi[k] = static_cast<float>(__ldg(P.pt_Vinv + 6 * static_cast<size_t>(p) + k));
            make_WY(Jca, Jpa, vi, W, Y);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float w0 = Jcb[j] * Jpb[0] + Jcb[6 + j] * Jpb[3];
                const float w1 = Jcb[j] * Jpb[1] + Jcb[6 + j] * Jpb[4];
                const float w2 = Jcb[j] * Jpb[2] + Jcb[6 + j] * Jpb[5];
#pragma unroll
                for (int i = 0; i < 6; ++i) acc[6 * i + j] -= Y[3 * i] * w0 + Y[3 * i + 1] * w1 + Y[3 * i + 2] * w2;
            }
        }
#pragma unroll
        for (int k = 0; k < 36; ++k) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) part[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < 36) {
            const int k = threadIdx.x;
            float v = part[0][k];
#pragma unroll
            for (int w = 1; w < kPairBlockThreads / 32; ++w) v += part[w][k];
            const int fa = b / P.n_free, fb = b - fa * P.n_free;
            sys[(static_cast<size_t>(fa) * 6 + k / 6) * n6 + static_cast<size_t>(fb) * 6 + (k % 6)] = static_cast<double>(v);
        }
        __syncthreads();
    }
}

// list of the camera-pair blocks that have tuples (order irrelevant: one CTA owns one block)
__global__ void compact_blocks_kernel(const int32_t* __restrict__ blk_start, long long nblk, int32_t* __restrict__ list,
                                      int32_t* __restrict__ counter) {
    for (long long b = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; b < nblk;
         b += static_cast<long long>(gridDim.x) * blockDim.x)
        if (blk_start[b + 1] > blk_start[b]) list[atomicAdd(counter, 1)] = static_cast<int32_t>(b);
}

// ---- structure building (once per problem)
// count / fill the co-observation tuples of every camera pair; mode 0 counts, mode 1 fills using cursors
__global__ void pair_tuples_kernel(Problem P, int mode, int32_t* __restrict__ blk_count_or_cursor, int2* __restrict__ tuples_out) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n_pts; p += gridDim.x * blockDim.x) {
        const int beg = P.pt_start[p], end = P.pt_start[p + 1];
        for (int a = beg; a < end; ++a) {
            const int fa = P.cam_free[P.obs_cam[a]];
            if (fa < 0) continue;
            for (int b = beg; b < end; ++b) {
                if (b == a) continue;
                const int fb = P.cam_free[P.obs_cam[b]];
                if (fb < 0 || !(fa < fb)) continue;          // upper block triangle; same-camera pairs are not expected
                const long long blk = static_cast<long long>(fa) * P.n_free + fb;
                if (mode == 0) atomicAdd(&blk_count_or_cursor[blk], 1);
                else {
                    const int slot = atomicAdd(&blk_count_or_cursor[blk], 1);
                    tuples_out[slot] = make_int2(a, b);
                }
            }
        }
    }
}
// single-block exclusive scan of int32 counts (n up to a few million) into out[0..n]; also copies the starts into cursor[]
__global__ void __launch_bounds__(1024) scan_i32_kernel(const int32_t* __restrict__ counts, long long n, int32_t* __restrict__ out,
                                                        int32_t* __restrict__ cursor) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < n; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const int v = (i < n) ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int ws = warp_sums[lane];
            int wincl = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wincl, o); if (lane >= o) wincl += y; }
            warp_sums[lane] = wincl - ws;
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - v;
        if (i < n) { out[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// S_ii += max(udiag_i, 1e-6) / radius  (after the all-reduce), and mirror nothing: the solver reads one triangle.
__global__ void damp_diagonal_kernel(double* __restrict__ sys, int n6, double inv_radius) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n6) return;
    const double* udiag = sys + static_cast<size_t>(n6) * n6 + 2 * static_cast<size_t>(n6);
    sys[static_cast<size_t>(i) * n6 + i] += fmax(udiag[i], 1e-6) * inv_radius;
}

// ------------------------------------------------------------------------------------------------ back-substitution
// dp = -V^-1 (g_p + sum_obs W^T dc);  candidate point = point + dp;  accumulates
// out[0] += model decrease  -(r.Jd + 1/2 |Jd|^2),  out[1] += |dp|^2,  out[2] += |point|^2
__global__ void __launch_bounds__(256)
backsub_kernel(Problem P, double inv_radius, const double* __restrict__ dc /*[n_free*6]*/, double* __restrict__ pts_new,
               double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    double acc_model = 0.0, acc_dp = 0.0, acc_x = 0.0;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < P.n_pts; p += gridDim.x * wpb) {
        const int beg = P.pt_start[p], end = P.pt_start[p + 1];
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        if (beg == end) {
            if (lane == 0) { pts_new[3 * p] = X[0]; pts_new[3 * p + 1] = X[1]; pts_new[3 * p + 2] = X[2]; }
            continue;
        }
        double Vinv[6], gp[3];
        LaneObs A;
        point_pass1(P, p, beg, end, lane, X, inv_radius, Vinv, gp, A);
        // shared focal block: dc holds (d fx, d fy) behind the camera steps
        const bool focal = P.refine_focal != 0;
        const double df0 = focal ? dc[static_cast<size_t>(P.n_free) * 6] : 0.0, df1 = focal ? dc[static_cast<size_t>(P.n_free) * 6 + 1] : 0.0;
        // t = g_p + sum W^T dc (+ Wf^T df) = g_p + sum Jp^T (Jc dc + Jf df)
        double t[3] = {0, 0, 0};
        for (int base = beg; base < end; base += 32) {
            if (base != beg) load_lane(P, base + lane, base + lane < end, X, A);
            if (A.valid && A.f >= 0) {
                double jd0 = 0.0, jd1 = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double d = dc[A.f * 6 + i];
                    jd0 += A.Jc[i] * d; jd1 += A.Jc[6 + i] * d;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] += A.Jp[k] * jd0 + A.Jp[3 + k] * jd1;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = gp[k] + warp_sum(t[k]);
        if (focal) {
            const double* wf = P.pt_Wf + 6 * static_cast<size_t>(p);
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] += wf[k] * df0 + wf[3 + k] * df1;
        }
        double dp[3];
        dp[0] = -(Vinv[0] * t[0] + Vinv[1] * t[1] + Vinv[2] * t[2]);
        dp[1] = -(Vinv[1] * t[0] + Vinv[3] * t[1] + Vinv[4] * t[2]);
        dp[2] = -(Vinv[2] * t[0] + Vinv[4] * t[1] + Vinv[5] * t[2]);
        if (lane == 0) {
            pts_new[3 * p] = X[0] + dp[0]; pts_new[3 * p + 1] = X[1] + dp[1]; pts_new[3 * p + 2] = X[2] + dp[2];
            acc_dp += dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
            acc_x += X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
        }
        // model decrease over the track
        for (int base = beg; base < end; base += 32) {
            if (!(base == beg && end - beg <= 32)) load_lane(P, base + lane, base + lane < end, X, A);
            if (A.valid) {
                double jd0 = A.Jp[0] * dp[0] + A.Jp[1] * dp[1] + A.Jp[2] * dp[2] + A.xy[0] * df0;
                double jd1 = A.Jp[3] * dp[0] + A.Jp[4] * dp[1] + A.Jp[5] * dp[2] + A.xy[1] * df1;
                if (A.f >= 0) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const double d = dc[A.f * 6 + i];
                        jd0 += A.Jc[i] * d; jd1 += A.Jc[6 + i] * d;
                    }
                }
                acc_model -= A.r[0] * jd0 + A.r[1] * jd1 + 0.5 * (jd0 * jd0 + jd1 * jd1);
            }
        }
    }
    __shared__ double sh[3][8];
    acc_model = warp_sum(acc_model); acc_dp = warp_sum(acc_dp); acc_x = warp_sum(acc_x);
    if (lane == 0) { sh[0][threadIdx.x >> 5] = acc_model; sh[1][threadIdx.x >> 5] = acc_dp; sh[2][threadIdx.x >> 5] = acc_x; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int k = 0; k < wpb; ++k) s += sh[threadIdx.x][k];
        atomicAdd(&out[threadIdx.x], s);
    }
}

// cams_new = cams + dc on the free cameras
__global__ void update_cams_kernel(const double* __restrict__ cams, const int32_t* __restrict__ cam_free, int n_cams,
                                   const double* __restrict__ dc, double* __restrict__ cams_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cams * 6) return;
    const int c = i / 6, k = i - 6 * c;
    const int f = cam_free[c];
    cams_new[i] = cams[i] + (f >= 0 ? dc[f * 6 + k] : 0.0);
}

// rhs -> double column used by the solver;  also negative step bookkeeping helpers
__global__ void copy_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

}  // namespace ba

// ------------------------------------------------------------------------------------------------ launchers
using namespace ba;
cudaError_t ba_launch_cam_prep(const double* cams, int n_cams, CamPre* pre, cudaStream_t st) {
    if (n_cams <= 0) return cudaSuccess;
    cam_prep_kernel<<<(n_cams + 127) / 128, 128, 0, st>>>(cams, n_cams, pre);
    return cudaGetLastError();
}
cudaError_t ba_launch_evaluate(const Problem& P, double* r_out, float* J_out, double* cost, int num_sms, cudaStream_t st) {
    if (P.n_obs <= 0) return cudaSuccess;
    int grid = (P.n_obs + 255) / 256;
    if (grid > num_sms * 8) grid = num_sms * 8;
    evaluate_kernel<<<grid, 256, 0, st>>>(P.pre, P.pts, P.obs_uv, P.obs_cam, P.obs_pt, P.n_obs, P.fx, P.fy, r_out, J_out, cost);
    return cudaGetLastError();
}
cudaError_t ba_launch_track_errors(const Problem& P, double* err, int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    int grid = (P.n_pts + 7) / 8;
    if (grid > num_sms * 8) grid = num_sms * 8;
    track_error_kernel<<<grid, 256, 0, st>>>(P, err);
    return cudaGetLastError();
}
cudaError_t ba_launch_linearize(const Problem& P, double inv_radius, double* sys, int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    int grid = (P.n_pts + 7) / 8;
    if (grid > num_sms * 8) grid = num_sms * 8;
    if (P.refine_focal) point_pass_kernel<true><<<grid, 256, 0, st>>>(P, inv_radius, sys);
    else point_pass_kernel<false><<<grid, 256, 0, st>>>(P, inv_radius, sys);
    if (P.n_free > 0) {
        camera_diag_kernel<<<P.n_free, 256, 0, st>>>(P, sys);
        if (P.refine_focal) camera_focal_border_kernel<<<P.n_free, 256, 0, st>>>(P, sys);
        if (P.n_blk_list > 0) {
            const int g2 = P.n_blk_list < num_sms * 64 ? P.n_blk_list : num_sms * 64;
            pair_block_kernel<<<g2, kPairBlockThreads, 0, st>>>(P, sys);
        }
    }
    return cudaGetLastError();
}
// structure lists: counts -> starts (+cursor) -> tuples.  blk_start has n_free^2 + 1 entries.
cudaError_t ba_launch_count_tuples(const Problem& P, int32_t* counts, int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0 || P.n_free <= 0) return cudaSuccess;
    pair_tuples_kernel<<<num_sms * 4, 256, 0, st>>>(P, 0, counts, nullptr);
    return cudaGetLastError();
}
cudaError_t ba_launch_scan_tuples(const int32_t* counts, long long n, int32_t* starts, int32_t* cursor, cudaStream_t st) {
    scan_i32_kernel<<<1, 1024, 0, st>>>(counts, n, starts, cursor);
    return cudaGetLastError();
}
cudaError_t ba_launch_fill_tuples(const Problem& P, int32_t* cursor, int2* tuples, int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0 || P.n_free <= 0) return cudaSuccess;
    pair_tuples_kernel<<<num_sms * 4, 256, 0, st>>>(P, 1, cursor, tuples);
    return cudaGetLastError();
}
cudaError_t ba_launch_compact_blocks(const int32_t* blk_start, long long nblk, int32_t* list, int32_t* counter, int num_sms,
                                     cudaStream_t st) {
    if (nblk <= 0) return cudaSuccess;
    compact_blocks_kernel<<<num_sms * 4, 256, 0, st>>>(blk_start, nblk, list, counter);
    return cudaGetLastError();
}
cudaError_t ba_launch_damp(double* sys, int n6, double inv_radius, cudaStream_t st) {
    if (n6 <= 0) return cudaSuccess;
    damp_diagonal_kernel<<<(n6 + 127) / 128, 128, 0, st>>>(sys, n6, inv_radius);
    return cudaGetLastError();
}
cudaError_t ba_launch_backsub(const Problem& P, double inv_radius, const double* dc, double* pts_new, double* out,
                              int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    int grid = (P.n_pts + 7) / 8;
    if (grid > num_sms * 8) grid = num_sms * 8;
    backsub_kernel<<<grid, 256, 0, st>>>(P, inv_radius, dc, pts_new, out);
    return cudaGetLastError();
}
cudaError_t ba_launch_update_cams(const double* cams, const int32_t* cam_free, int n_cams, const double* dc,
                                  double* cams_new, cudaStream_t st) {
    if (n_cams <= 0) return cudaSuccess;
    update_cams_kernel<<<(n_cams * 6 + 127) / 128, 128, 0, st>>>(cams, cam_free, n_cams, dc, cams_new);
    return cudaGetLastError();
}
cudaError_t ba_launch_copy(const double* src, double* dst, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
    return cudaGetLastError();
}

}  // namespace msfm
