// K2 — bundle-adjustment inner loop on the device.
//
// Replaces what Ceres does behind CeresBundelOptimizer::Optimize (src/Optimizer/CeresBundleOptimizer.cpp:293):
//   * per observation: the residual of BundleAutoDiffConstantFocalCostFunction (:29-53) and its 2 x (3+3+3)
//     Jacobian — here ANALYTIC, equal to the autodiff of Ceres' AngleAxisRotatePoint in both of its branches;
//   * per point: V = sum Jp^T Jp (+ Marquardt damping), g_p; per camera: U, g_c;
//   * Schur elimination of the points onto the free cameras: S = U + D_c - sum W V^-1 W^T, rhs = -(g_c - W V^-1 g_p)
//     (Ceres SchurEliminator for DENSE_SCHUR / SPARSE_SCHUR, :264-273);
//   * back-substitution of the points, candidate-step evaluation.
//
// ONE kernel per linearisation (fused_linearize_kernel; a small pre-pass only for tracks longer than 32 views).  Nothing
// per-observation is written to memory: the 2 x 6 / 2 x 3 Jacobians of a TILE's observations (ba_types.cuh: Tile, at most
// 512 observations over at most 32 cameras) are staged in shared memory, the 6x6 products
// Y_a W_b^T = Jc_a^T (Jp_a V^-1 Jp_b^T) Jc_b of all camera pairs of the tile's points are accumulated into the CTA's
// shared-memory copy of the blocks the tile touches, and that copy is flushed once per tile with vector reductions
// (red.global.add.v4.f32) into the block-sparse S.  History: v1 scattered every product with
// fp64 global atomics (bound by L2 atomic throughput, profiles/r01_k2_ncu_raw.csv); v2 stored Jacobians and residuals
// per observation and gathered them again per camera and per camera pair (three kernels, ~10x the algorithmic DRAM
// traffic, latency-bound gathers: profiles/r01b_k2_ncu_summary.txt).
// Geometry (projection, residual, V^-1, gradients) is evaluated in fp64, the 6x6 block products in fp32.  Tensor
// cores are not used: sparse, irregular work.
#include <cuda_runtime.h>
#include <cstdint>

#include "ba_types.cuh"
#include "launch_count.hpp"

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
    p.pad = 0.0;
    pre[c] = p;
}

// ------------------------------------------------------------------------------------------------ per observation
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

// residual + Jacobian dump for parity tests (msfm_ba_evaluate); outputs in the CALLER's observation order
__global__ void evaluate_kernel(Problem P, double* __restrict__ r_out, float* __restrict__ J_out, double* __restrict__ cost) {
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_obs; i += gridDim.x * blockDim.x) {
        const CamPre c = P.pre[P.obs_cam[i]];
        const int p = P.obs_pt[i];
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        double r[2], Jc[12], Jp[6];
        if (J_out) obs_eval<true>(c, X, P.obs_uv[2 * i], P.obs_uv[2 * i + 1], P.fx, P.fy, r, Jc, Jp);
        else obs_eval<false>(c, X, P.obs_uv[2 * i], P.obs_uv[2 * i + 1], P.fx, P.fy, r, Jc, Jp);
        local += 0.5 * (r[0] * r[0] + r[1] * r[1]);
        const size_t o = (r_out || J_out) ? static_cast<size_t>(P.obs_orig[i]) : 0;
        if (r_out) { r_out[2 * o] = r[0]; r_out[2 * o + 1] = r[1]; }
        if (J_out) {
            float* J = J_out + o * 18;
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
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Mean reprojection error of every point over its observations, sqrt(dx^2 + dy^2) per observation: what
// Map::UpdateFromBAData recomputes on the host after every BA through ComputeTrackError
// (src/Reconstruction/Map.cpp:1201, 1834-1846; Projection::CalculateReprojectionError, Projection.cpp:114-133).
// One warp per point, one lane per observation.
__global__ void __launch_bounds__(256)
track_error_kernel(Problem P, double* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int d = blockIdx.x * wpb + (threadIdx.x >> 5); d < P.n_pts; d += gridDim.x * wpb) {
        const int p = P.pt_order[d];
        const int beg = P.pt_start[d], end = P.pt_start[d + 1];
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

// Post-BA filter statistics: what Map::FilterAllPoints3D recomputes on the host after every global BA
// (src/Reconstruction/Map.cpp:793-917): per observation HasPositiveDepth && reprojection error <= max_reproj_error
// (Projection.cpp:6-19, 114-133), per point the mean error over the observations that pass and the LARGEST parallax angle
// over its camera pairs (Projection::CalculateParallaxAngle, Projection.cpp:149-194: law of cosines on the rays from the two
// projection centres O = -R^T t, min(angle, pi - angle), degrees, NaN -> 0).  One warp per point, one lane per observation.
__global__ void __launch_bounds__(256)
filter_stats_kernel(Problem P, double max_reproj_error, uint8_t* __restrict__ obs_keep, double* __restrict__ pt_err,
                    int32_t* __restrict__ pt_kept, double* __restrict__ pt_angle) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int d = blockIdx.x * wpb + (threadIdx.x >> 5); d < P.n_pts; d += gridDim.x * wpb) {
        const int p = P.pt_order[d];
        const int beg = P.pt_start[d], end = P.pt_start[d + 1];
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        double sum = 0.0, amax = 0.0;
        int kept = 0;
        for (int o = beg + lane; o < end; o += 32) {
            const CamPre c = P.pre[P.obs_cam[o]];
            const double pz = c.R[6] * X[0] + c.R[7] * X[1] + c.R[8] * X[2] + c.t[2];
            double r[2], Jc[12], Jp[6];
            obs_eval<false>(c, X, P.obs_uv[2 * o], P.obs_uv[2 * o + 1], P.fx, P.fy, r, Jc, Jp);
            const double e = sqrt(r[0] * r[0] + r[1] * r[1]);
            const bool keep = pz > 0.0 && !(e > max_reproj_error);
            if (obs_keep) obs_keep[P.obs_orig[o]] = keep ? 1 : 0;
            if (keep) { sum += e; kept += 1; }
            // parallax against every earlier observation of the track
            const double O1[3] = {-(c.R[0] * c.t[0] + c.R[3] * c.t[1] + c.R[6] * c.t[2]), -(c.R[1] * c.t[0] + c.R[4] * c.t[1] + c.R[7] * c.t[2]),
                                  -(c.R[2] * c.t[0] + c.R[5] * c.t[1] + c.R[8] * c.t[2])};
            const double ray1 = sqrt((X[0] - O1[0]) * (X[0] - O1[0]) + (X[1] - O1[1]) * (X[1] - O1[1]) + (X[2] - O1[2]) * (X[2] - O1[2]));
            for (int q = beg; q < o; ++q) {
                const CamPre& c2 = P.pre[P.obs_cam[q]];
                const double O2[3] = {-(c2.R[0] * c2.t[0] + c2.R[3] * c2.t[1] + c2.R[6] * c2.t[2]), -(c2.R[1] * c2.t[0] + c2.R[4] * c2.t[1] + c2.R[7] * c2.t[2]),
                                      -(c2.R[2] * c2.t[0] + c2.R[5] * c2.t[1] + c2.R[8] * c2.t[2])};
                const double ray2 = sqrt((X[0] - O2[0]) * (X[0] - O2[0]) + (X[1] - O2[1]) * (X[1] - O2[1]) + (X[2] - O2[2]) * (X[2] - O2[2]));
                const double base2 = (O1[0] - O2[0]) * (O1[0] - O2[0]) + (O1[1] - O2[1]) * (O1[1] - O2[1]) + (O1[2] - O2[2]) * (O1[2] - O2[2]);
                const double ang = fabs(acos((ray1 * ray1 + ray2 * ray2 - base2) / (2.0 * ray1 * ray2)));
                const double deg = isnan(ang) ? 0.0 : fmin(ang, 3.14159265358979323846 - ang) * 180.0 / 3.14159265358979323846;
                amax = fmax(amax, deg);
            }
        }
        sum = warp_sum(sum);
        amax = warp_max(amax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
        if (lane == 0) {
            if (pt_err) pt_err[p] = kept > 0 ? sum / static_cast<double>(kept) : 0.0;
            if (pt_kept) pt_kept[p] = kept;
            if (pt_angle) pt_angle[p] = amax;
        }
    }
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
    o.cam = __ldg(P.obs_cam + obs);
    o.f = __ldg(P.cam_free + o.cam);
    const CamPre c = P.pre[o.cam];
    const double2 uv = __ldg(reinterpret_cast<const double2*>(P.obs_uv) + obs);
    double Jc[12], Jp[6];
    obs_eval<true>(c, X, uv.x, uv.y, P.fx, P.fy, o.r, Jc, Jp, o.xy);
    if (o.f >= 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) o.Jc[k] = static_cast<float>(Jc[k]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) o.Jp[k] = static_cast<float>(Jp[k]);
}

// per-point normal-equation pieces: V (damped), V^-1, g_p — reduced over the whole track (observations beg..end-1 in
// chunks of 32, one per lane).  `first` keeps the lane's observation of the first chunk; cost_lane adds the lane's share
// of 1/2 sum r^2.
// kFocal: also wf = Wf = sum Jf^T Jp (2x3) and fstat = sum xp^2, sum yp^2, sum xp r0, sum yp r1 over the track.
// A template parameter, not a run-time flag: the extra accumulators cost the constant-focal kernel 70 registers when they
// were merely branched around.
template <bool kFocal = false>
__device__ __forceinline__ void point_pass1(const Problem& P, int beg, int end, int lane, const double X[3],
                                            double inv_radius, double Vinv[6], double gp[3], LaneObs& first,
                                            double& cost_lane, double* wf = nullptr, double* fstat = nullptr) {
    double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    double w6[6] = {0, 0, 0, 0, 0, 0}, f4[4] = {0, 0, 0, 0};
    for (int base = beg; base < end; base += 32) {
        LaneObs o;
        load_lane(P, base + lane, base + lane < end, X, o);
        if (base == beg) first = o;
        cost_lane += 0.5 * (o.r[0] * o.r[0] + o.r[1] * o.r[1]);
        const double j0 = o.Jp[0], j1 = o.Jp[1], j2 = o.Jp[2], j3 = o.Jp[3], j4 = o.Jp[4], j5 = o.Jp[5];
        v[0] += j0 * j0 + j3 * j3; v[1] += j0 * j1 + j3 * j4; v[2] += j0 * j2 + j3 * j5;
        v[3] += j1 * j1 + j4 * j4; v[4] += j1 * j2 + j4 * j5; v[5] += j2 * j2 + j5 * j5;
        g[0] += j0 * o.r[0] + j3 * o.r[1]; g[1] += j1 * o.r[0] + j4 * o.r[1]; g[2] += j2 * o.r[0] + j5 * o.r[1];
        if (kFocal) {
            w6[0] += o.xy[0] * j0; w6[1] += o.xy[0] * j1; w6[2] += o.xy[0] * j2;
            w6[3] += o.xy[1] * j3; w6[4] += o.xy[1] * j4; w6[5] += o.xy[1] * j5;
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
// Shared-memory row of one staged observation (floats): Jc[12] | Q[6] | Jp[6] | local camera (-1: constant pose) | pad[3].
// 28 floats = 7 x 16 bytes: 128-bit loads; rows r and r' collide on a bank group only when r = r' mod 8.
constexpr int kStageStride = 28;
constexpr int kFusedThreads = kTileObs;       // one thread per observation of a tile
constexpr int kTileRunItems = 2 * kTileObs;           // work items of the pair phase staged in shared memory (more: read from global)

// acc[0..5] += v[0..5] on shared memory, 8-byte aligned.  sm_100 has no native fp32 shared-memory atomic add (atomicAdd
// compiles to an LDS / FADD / ATOMS.CAST.SPIN loop per element, one after the other); here a row of a block is three 64-bit
// compare-and-swaps in flight together, with ONE branch for the rare lost race.
__device__ __forceinline__ void smem_add_row(float* base, const float v[6]) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(base);
    const float2 o0 = *reinterpret_cast<const float2*>(base), o1 = *reinterpret_cast<const float2*>(base + 2),
                 o2 = *reinterpret_cast<const float2*>(base + 4);
    const unsigned long long e0 = (static_cast<unsigned long long>(__float_as_uint(o0.y)) << 32) | __float_as_uint(o0.x);
    const unsigned long long e1 = (static_cast<unsigned long long>(__float_as_uint(o1.y)) << 32) | __float_as_uint(o1.x);
    const unsigned long long e2 = (static_cast<unsigned long long>(__float_as_uint(o2.y)) << 32) | __float_as_uint(o2.x);
    const unsigned long long n0 = (static_cast<unsigned long long>(__float_as_uint(o0.y + v[1])) << 32) | __float_as_uint(o0.x + v[0]);
    const unsigned long long n1 = (static_cast<unsigned long long>(__float_as_uint(o1.y + v[3])) << 32) | __float_as_uint(o1.x + v[2]);
    const unsigned long long n2 = (static_cast<unsigned long long>(__float_as_uint(o2.y + v[5])) << 32) | __float_as_uint(o2.x + v[4]);
    const unsigned long long g0 = atomicCAS(b, e0, n0), g1 = atomicCAS(b + 1, e1, n1), g2 = atomicCAS(b + 2, e2, n2);
    if ((g0 != e0) | (g1 != e1) | (g2 != e2)) {
        if (g0 != e0) { atomicAdd(base, v[0]); atomicAdd(base + 1, v[1]); }
        if (g1 != e1) { atomicAdd(base + 2, v[2]); atomicAdd(base + 3, v[3]); }
        if (g2 != e2) { atomicAdd(base + 4, v[4]); atomicAdd(base + 5, v[5]); }
    }
}
// blk[0..35] += v (18 float2), blk 16-byte aligned: nine 128-bit compare-and-swaps (ATOMS.CAS.128), three in flight at a
// time, one branch per three for the rare lost race.
__device__ __forceinline__ void smem_add_block(float* blk, const float2 v[18]) {
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(blk));
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        unsigned long long elo[3], ehi[3], glo[3], ghi[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int q = 3 * g + t;
            const float4 o = reinterpret_cast<const float4*>(blk)[q];
            const float2 n0 = __fadd2_rn(make_float2(o.x, o.y), v[2 * q]), n1 = __fadd2_rn(make_float2(o.z, o.w), v[2 * q + 1]);
            elo[t] = (static_cast<unsigned long long>(__float_as_uint(o.y)) << 32) | __float_as_uint(o.x);
            ehi[t] = (static_cast<unsigned long long>(__float_as_uint(o.w)) << 32) | __float_as_uint(o.z);
            const unsigned long long nlo = (static_cast<unsigned long long>(__float_as_uint(n0.y)) << 32) | __float_as_uint(n0.x);
            const unsigned long long nhi = (static_cast<unsigned long long>(__float_as_uint(n1.y)) << 32) | __float_as_uint(n1.x);
            asm volatile("{\n .reg .b128 e, n, g;\n mov.b128 e, {%2, %3};\n mov.b128 n, {%4, %5};\n atom.shared.cas.b128 g, [%6], e, n;\n mov.b128 {%0, %1}, g;\n}"
                         : "=l"(glo[t]), "=l"(ghi[t]) : "l"(elo[t]), "l"(ehi[t]), "l"(nlo), "l"(nhi), "r"(base + 16u * q) : "memory");
        }
        bool lost = false;
#pragma unroll
        for (int t = 0; t < 3; ++t) lost |= (glo[t] != elo[t]) | (ghi[t] != ehi[t]);
        if (lost) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if ((glo[t] != elo[t]) | (ghi[t] != ehi[t])) {
                    float* d = blk + 4 * (3 * g + t);
                    atomicAdd(d, v[2 * (3 * g + t)].x); atomicAdd(d + 1, v[2 * (3 * g + t)].y);
                    atomicAdd(d + 2, v[2 * (3 * g + t) + 1].x); atomicAdd(d + 3, v[2 * (3 * g + t) + 1].y);
                }
        }
    }
}
__device__ __forceinline__ void smem_add6(double* base, const double v[6]) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(base);
    double old[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) old[e] = *reinterpret_cast<volatile double*>(base + e);
    unsigned long long got[6];
    bool lost = false;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        got[e] = atomicCAS(b + e, static_cast<unsigned long long>(__double_as_longlong(old[e])),
                           static_cast<unsigned long long>(__double_as_longlong(old[e] + v[e])));
        lost |= got[e] != static_cast<unsigned long long>(__double_as_longlong(old[e]));
    }
    if (lost) {
#pragma unroll
        for (int e = 0; e < 6; ++e)
            if (got[e] != static_cast<unsigned long long>(__double_as_longlong(old[e]))) atomicAdd(base + e, v[e]);
    }
}

// Rows I0 and 5 - I0 of a camera's diagonal block U - Y W^T = sum Jc^T N Jc over the observations cam_obs[pb], cam_obs[pb + step],
// ...: UPPER triangle only (the block is symmetric: 21 of its 36 entries; every reader takes the upper triangle), 7 sums.
template <int I0>
__device__ __forceinline__ void diag_row_pair(const float* __restrict__ stage, const uint16_t* __restrict__ cam_obs, int pb, int pe, int step,
                                              float* blk) {
    constexpr int I1 = 5 - I0;
    float r0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pp = pb; pp < pe; pp += step) {
        const float* sr = stage + static_cast<int>(cam_obs[pp]) * kStageStride;
        const float4 j0 = reinterpret_cast<const float4*>(sr)[0], j1 = reinterpret_cast<const float4*>(sr)[1],
                     j2 = reinterpret_cast<const float4*>(sr)[2], nn = reinterpret_cast<const float4*>(sr)[6];
        const float n00 = nn.y, n01 = nn.z, n11 = nn.w;
        const float c0[6] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y}, c1[6] = {j1.z, j1.w, j2.x, j2.y, j2.z, j2.w};
        const float ua = c0[I0] * n00 + c1[I0] * n01, va = c0[I0] * n01 + c1[I0] * n11;     // row I0 of Jc^T N
        const float ub = c0[I1] * n00 + c1[I1] * n01, vb = c0[I1] * n01 + c1[I1] * n11;     // row I1 of Jc^T N
#pragma unroll
        for (int j = I0; j < 6; ++j) r0[j] += ua * c0[j] + va * c1[j];
#pragma unroll
        for (int j = I1; j < 6; ++j) r1[j] += ub * c0[j] + vb * c1[j];
    }
    if (pb < pe) {
        smem_add_row(blk + 6 * I0, r0);
        smem_add_row(blk + 6 * I1, r1);
    }
}

// Everything a CTA needs to know about a tile before it touches an observation, staged in shared memory one tile AHEAD (two
// buffers): the Tile record, its local cameras, their free-camera indices and the block slots of its local camera pairs.
// (Read from global memory at the point of use these were a chain of dependent L2 round trips at the top of every tile and one
// exposed load per flushed block row: 12 % of the kernel's stall samples at `if (slot < 0)` alone, profiles/r02_k2_v6_source_lines.txt.)
struct alignas(16) TileHdr {
    Tile T;
    int32_t cams[kTileCams];
    int32_t free_[kTileCams];
    int32_t slots[kTileCams * (kTileCams + 1) / 2];
};
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Shared-memory carve-up of the linearisation kernel
struct FusedSmem {
    size_t camacc, acc, stage, rbuf, qbuf, xybuf, ptV, ptg, ptWf, cam_start, cam_cursor, cam_obs, unit_info, run_items, hdr, misc, total;
};
__host__ __device__ inline FusedSmem fused_smem_layout(bool focal) {
    FusedSmem L;
    size_t o = 0;
    L.camacc = o; o += static_cast<size_t>(kTileCams) * (focal ? 30 : 18) * sizeof(double);
    L.ptg = o; o += static_cast<size_t>(kTilePts) * 3 * sizeof(double);
    L.rbuf = o; o += static_cast<size_t>(kTileObs) * 2 * sizeof(double);
    L.qbuf = o; o += static_cast<size_t>(kTileObs) * 2 * sizeof(double);
    L.xybuf = o; o += focal ? static_cast<size_t>(kTileObs) * 2 * sizeof(double) : 0;
    L.ptWf = o; o += focal ? static_cast<size_t>(kTilePts) * 6 * sizeof(double) : 0;
    L.acc = o; o += static_cast<size_t>(kTileCams) * (kTileCams + 1) / 2 * kBlkStride * sizeof(float);
    L.stage = o; o += static_cast<size_t>(kTileObs) * kStageStride * sizeof(float);
    L.ptV = o; o += static_cast<size_t>(kTilePts) * 6 * sizeof(float);
    L.cam_start = o; o += (static_cast<size_t>(kTileCams) + 4) * sizeof(int32_t);
    L.cam_cursor = o; o += static_cast<size_t>(kTileCams) * sizeof(int32_t);
    L.cam_obs = o; o += static_cast<size_t>(kTileObs) * sizeof(uint16_t);
    L.unit_info = o; o += static_cast<size_t>(kTilePts) * sizeof(uint32_t);
    L.run_items = o; o += static_cast<size_t>(kTileRunItems) * sizeof(uint32_t);
    o = (o + 15) / 16 * 16;
    L.hdr = o; o += 2 * sizeof(TileHdr);
    L.misc = o; o += 64;
    L.total = (o + 15) / 16 * 16;
    return L;
}

// Pre-pass over the long tracks (more than 32 observations): V^-1, g_p (and Wf) of every long point, its share of the
// cost / gradient maximum / focal sums.  The items of the point read these instead of reducing over the whole track again.
template <bool kFocal>
__global__ void __launch_bounds__(256)
long_track_prepass_kernel(Problem P, double inv_radius) {
    // warp index / track bounds through a shuffle: provably warp-uniform, so the reductions' shuffles stay plain SHFL (backsub_kernel)
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    constexpr int kRec = kFocal ? 15 : 9;
    double cost_local = 0.0, gpmax_local = 0.0;
    double ff[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int li = blockIdx.x * wpb + warp; li < P.n_long; li += gridDim.x * wpb) {
        const int d = P.first_long + li;
        const int p = __ldg(P.pt_order + d);
        const int beg = __shfl_sync(0xffffffffu, __ldg(P.pt_start + d), 0), end = __shfl_sync(0xffffffffu, __ldg(P.pt_start + d + 1), 0);
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        double Vinv[6], gp[3], Wf[6], fs[4], cl = 0.0;
        LaneObs A;
        point_pass1<kFocal>(P, beg, end, lane, X, inv_radius, Vinv, gp, A, cl, Wf, fs);
        cost_local += cl;
        gpmax_local = fmax(gpmax_local, fmax(fabs(gp[0]), fmax(fabs(gp[1]), fabs(gp[2]))));
        double* rec = P.long_V + static_cast<size_t>(li) * kRec;
        if (lane < 6) rec[lane] = Vinv[lane];
        if (lane < 3) rec[6 + lane] = gp[lane];
        if (kFocal) {
            if (lane < 6) { rec[9 + lane] = Wf[lane]; P.pt_Wf[6 * static_cast<size_t>(d) + lane] = Wf[lane]; }
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
    }
    __shared__ double sh_c[8], sh_g[8], sh_f[8][9];
    cost_local = warp_sum(cost_local);                    // lanes hold disjoint shares of the cost
    gpmax_local = warp_max(gpmax_local);
    if (lane == 0) {
        sh_c[warp] = cost_local; sh_g[warp] = gpmax_local;
#pragma unroll
        for (int k = 0; k < 9; ++k) sh_f[warp][k] = ff[k];     // lane-uniform
    }
    __syncthreads();
    if (kFocal && threadIdx.x < 9) {
        double v = 0.0;
        for (int k = 0; k < wpb; ++k) v += sh_f[k][threadIdx.x];
        atomicAdd(P.tail + P.tl.ff + threadIdx.x, v);
    }
    if (threadIdx.x == 0) {
        double c = 0.0, g = 0.0;
        for (int k = 0; k < wpb; ++k) { c += sh_c[k]; g = fmax(g, sh_g[k]); }
        atomicAdd(P.tail + P.tl.scal, c);
        atomicMax(reinterpret_cast<unsigned long long*>(P.tail + P.tl.gpm + P.gpm_slot), static_cast<unsigned long long>(__double_as_longlong(g)));
    }
}

// One CTA per tile (dynamic scheduler), 512 threads.  Per tile:
//   A  thread = observation: projection, residual, Jacobians (fp64) -> staged in shared memory as fp32 (r stays fp64)
//   B  thread = point (unit): V, g_p over its staged observations, damping, 3x3 inverse
//   C  thread = observation: Q = Jp V^-1, N = I - Q Jp^T, q = Q g_p - r -> staged; observations bucketed by local camera
//   E  a work queue over the warps:
//        camera items  lane = local camera: ONE part (two rows of the upper triangle of the diagonal block U - Y W^T =
//                      sum Jc^T N Jc, or rhs / g_c / diag U) summed over a quarter of the camera's observations of the tile —
//                      four adders per address;
//        run items     lanes = 32 camera pairs (x < y) of a RUN of points with identical camera lists: block (x, y) -=
//                      sum over the run of Jc_x^T (Q_x Jp_y^T) Jc_y, summed in registers with packed fp32x2 FMAs and added to
//                      the shared-memory block with 128-bit compare-and-swaps (the pairs of a point never share a block;
//                      different warps work on different points, so lost races are rare)
//   flush: blocks with red.global.add.v4.f32, camera vectors with fp64 reductions.
// (v4 of this kernel let every observation thread add its diagonal contribution itself: 512 threads into 32 cameras' accumulators
// at the same instant, 16-way compare-and-swap contention, 86k cycles per tile — profiles/r02_k2_history.txt.)
template <bool kFocal>
__global__ void __launch_bounds__(kFusedThreads, kCtasPerSm)
fused_linearize_kernel(Problem P, double inv_radius) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CV = kFocal ? 30 : 18;         // per local camera: rhs[6] | gc[6] | udiag[6] (| border B [6][2])
    const FusedSmem L = fused_smem_layout(kFocal);
    double* camacc = reinterpret_cast<double*>(smem_raw + L.camacc);
    double* ptg = reinterpret_cast<double*>(smem_raw + L.ptg);
    double* rbuf = reinterpret_cast<double*>(smem_raw + L.rbuf);
    double* xybuf = reinterpret_cast<double*>(smem_raw + L.xybuf);
    double* ptWf = reinterpret_cast<double*>(smem_raw + L.ptWf);
    float* acc = reinterpret_cast<float*>(smem_raw + L.acc);
    float* stage = reinterpret_cast<float*>(smem_raw + L.stage);
    float* ptV = reinterpret_cast<float*>(smem_raw + L.ptV);
    double* qbuf = reinterpret_cast<double*>(smem_raw + L.qbuf);
    int32_t* cam_start = reinterpret_cast<int32_t*>(smem_raw + L.cam_start);      // [w + 1]; doubles as the histogram in phase A
    int32_t* cam_cursor = reinterpret_cast<int32_t*>(smem_raw + L.cam_cursor);
    uint16_t* cam_obs = reinterpret_cast<uint16_t*>(smem_raw + L.cam_obs);        // observation rows bucketed by local camera
    uint32_t* unit_info = reinterpret_cast<uint32_t*>(smem_raw + L.unit_info);       // obs base (16) | nA (8) | nB (8)
    uint32_t* run_items = reinterpret_cast<uint32_t*>(smem_raw + L.run_items);
    TileHdr* hdr = reinterpret_cast<TileHdr*>(smem_raw + L.hdr);
    int32_t* misc = reinterpret_cast<int32_t*>(smem_raw + L.misc);                   // [1] work-queue cursor, [4], [5] tile index of header 0 / 1
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* tail = P.tail;
    double cost_local = 0.0, gpmax_local = 0.0;
    double ff[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};        // focal block sums of this thread's points
    if (tid <= kTileCams) cam_start[tid] = 0;
    // ---- tile pipeline: header `cur` is complete when a tile starts; the Tile record of the next tile is copied while phases
    //      A - C run, its camera / slot tables while phase E runs; the index of the tile after next comes from an atomic whose
    //      latency hides behind phase A
    if (tid == 0) { misc[4] = atomicAdd(P.tile_counter, 1); misc[5] = atomicAdd(P.tile_counter, 1); }
    __syncthreads();
    {
        const int t0 = misc[4];
        if (t0 < P.n_tiles) {
            const Tile T0 = P.tiles[t0];
            if (tid == 0) hdr[0].T = T0;
            if (tid < T0.w) { hdr[0].cams[tid] = __ldg(P.tile_cams + T0.cam_begin + tid); hdr[0].free_[tid] = __ldg(P.tile_free + T0.cam_begin + tid); }
            for (int i = tid; i < T0.w * (T0.w + 1) / 2; i += kFusedThreads) hdr[0].slots[i] = __ldg(P.tile_slots + T0.slot_begin + i);
        }
    }
    __syncthreads();
    int cur = 0;

    for (;;) {
        const int ti = misc[4 + cur];
        if (ti >= P.n_tiles) break;
        const TileHdr& H = hdr[cur];
        const Tile T = H.T;
        const int ti_next = misc[4 + (cur ^ 1)];
        int ti_after = 0;
        if (tid == 0) { ti_after = atomicAdd(P.tile_counter, 1); misc[1] = 0; }
        if (tid < 3 && ti_next < P.n_tiles)
            cp_async16(reinterpret_cast<unsigned char*>(&hdr[cur ^ 1].T) + 16 * tid, reinterpret_cast<const unsigned char*>(P.tiles + ti_next) + 16 * tid);
        const bool split = (T.flags & kTileSplit) != 0;
        const int n_units = T.end - T.begin;
        const int nb = T.w * (T.w + 1) / 2;
        {
            float4* a4 = reinterpret_cast<float4*>(acc);
            for (int i = tid; i < nb * (kBlkStride / 4); i += kFusedThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < T.w * CV; i += kFusedThreads) camacc[i] = 0.0;
            for (int i = tid; i < T.n_runs && i < kTileRunItems; i += kFusedThreads) run_items[i] = __ldg(P.runs + T.run_begin + i);
        }
        // ---- A: this thread's observation
        int unit = 0, f_mine = -1, lcam = 0;
        bool valid, diag_owner;
        {
            int o, d;
            if (!split) {
                valid = tid < T.n_obs;
                o = T.obs_begin + tid;
                unit = valid ? static_cast<int>(__ldg(P.obs_lpt + o)) : 0;
                d = T.begin + unit;
                lcam = valid ? static_cast<int>(__ldg(P.obs_lcam + o)) : 0;
                diag_owner = true;
                if (tid < n_units) {
                    const int b0 = __ldg(P.pt_start + T.begin + tid), b1 = __ldg(P.pt_start + T.begin + tid + 1);
                    unit_info[tid] = static_cast<uint32_t>(b0 - T.obs_begin) | (static_cast<uint32_t>(b1 - b0) << 16);
                }
            } else {
                unit = tid >> 5;
                const bool live = unit < n_units;
                const Item* item = P.items + T.begin + (live ? unit : 0);
                const int a0 = item->a0, a1 = item->a1, b0 = item->b0, b1 = item->b1;
                const int nA = a1 - a0, nB = b1 - b0;
                valid = live && lane < nA + nB;
                d = item->d;
                o = __ldg(P.pt_start + d) + (lane < nA ? a0 + lane : b0 + lane - nA);
                lcam = item->lc[lane];
                diag_owner = nB == 0;
                if (live && lane == 0) unit_info[unit] = static_cast<uint32_t>(unit * 32) | (static_cast<uint32_t>(nA) << 16) | (static_cast<uint32_t>(nB) << 24);
            }
            float* srow = stage + tid * kStageStride;
            if (valid) {
                const int cam = H.cams[lcam];
                f_mine = H.free_[lcam];
                const int p = split ? __ldg(P.pt_order + d) : __ldg(P.obs_pt + o);      // the caller's point index
                const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
                const double2 uv = __ldg(reinterpret_cast<const double2*>(P.obs_uv) + o);
                double r[2], Jc[12], Jp[6], xy[2];
                obs_eval<true>(P.pre[cam], X, uv.x, uv.y, P.fx, P.fy, r, Jc, Jp, xy);
                float4* s4 = reinterpret_cast<float4*>(srow);
                if (f_mine >= 0) {
                    s4[0] = make_float4(static_cast<float>(Jc[0]), static_cast<float>(Jc[1]), static_cast<float>(Jc[2]), static_cast<float>(Jc[3]));
                    s4[1] = make_float4(static_cast<float>(Jc[4]), static_cast<float>(Jc[5]), static_cast<float>(Jc[6]), static_cast<float>(Jc[7]));
                    s4[2] = make_float4(static_cast<float>(Jc[8]), static_cast<float>(Jc[9]), static_cast<float>(Jc[10]), static_cast<float>(Jc[11]));
                } else {
                    s4[0] = s4[1] = s4[2] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                srow[18] = static_cast<float>(Jp[0]); srow[19] = static_cast<float>(Jp[1]); srow[20] = static_cast<float>(Jp[2]);
                srow[21] = static_cast<float>(Jp[3]); srow[22] = static_cast<float>(Jp[4]); srow[23] = static_cast<float>(Jp[5]);
                srow[24] = __int_as_float(f_mine >= 0 ? lcam : -1);
                if (f_mine >= 0 && diag_owner) atomicAdd(cam_start + 1 + lcam, 1);      // bucket sizes (native integer atomic)
                rbuf[2 * tid] = r[0]; rbuf[2 * tid + 1] = r[1];
                if (kFocal) { xybuf[2 * tid] = xy[0]; xybuf[2 * tid + 1] = xy[1]; }
            }
        }
        if (tid < 3) cp_async_wait_all();                       // the next tile's record has landed (visible after the barrier)
        __syncthreads();
        if (tid == 0) misc[4 + cur] = ti_after;                 // this header slot serves the tile after next
        // ---- B: this thread's point: V^-1, g_p
        {
            // sB threads per point (a power of two <= 8 that keeps all points of the tile inside the CTA): each sums every sB-th
            // observation, the partial sums meet through shuffles.  (One thread per point left 50 - 250 of the 512 threads with a
            // serial loop over the whole track while the others waited at the barrier: 9 % of the stall samples.)
            int sB = 1, lgB = 0;
            if (!split)
                while (sB < 8 && n_units * sB * 2 <= kFusedThreads) { sB *= 2; ++lgB; }
            const int u = tid >> lgB, sub = tid & (sB - 1);
            const bool mine = u < n_units;
            double Vinv[6], gp[3], Wf[6] = {0, 0, 0, 0, 0, 0};
            if (!split) {
                double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, f4[4] = {0, 0, 0, 0}, cl = 0.0;
                if (mine) {
                    const uint32_t ui = unit_info[u];
                    const int ob = static_cast<int>(ui & 0xFFFFu), k = static_cast<int>(ui >> 16);
                    for (int j = sub; j < k; j += sB) {
                        const float* sr = stage + (ob + j) * kStageStride;
                        const double j0 = sr[18], j1 = sr[19], j2 = sr[20], j3 = sr[21], j4 = sr[22], j5 = sr[23];
                        const double r0 = rbuf[2 * (ob + j)], r1 = rbuf[2 * (ob + j) + 1];
                        cl += 0.5 * (r0 * r0 + r1 * r1);
                        v[0] += j0 * j0 + j3 * j3; v[1] += j0 * j1 + j3 * j4; v[2] += j0 * j2 + j3 * j5;
                        v[3] += j1 * j1 + j4 * j4; v[4] += j1 * j2 + j4 * j5; v[5] += j2 * j2 + j5 * j5;
                        g[0] += j0 * r0 + j3 * r1; g[1] += j1 * r0 + j4 * r1; g[2] += j2 * r0 + j5 * r1;
                        if (kFocal) {
                            const double xp = xybuf[2 * (ob + j)], yp = xybuf[2 * (ob + j) + 1];
                            Wf[0] += xp * j0; Wf[1] += xp * j1; Wf[2] += xp * j2; Wf[3] += yp * j3; Wf[4] += yp * j4; Wf[5] += yp * j5;
                            f4[0] += xp * xp; f4[1] += yp * yp; f4[2] += xp * r0; f4[3] += yp * r1;
                        }
                    }
                }
                cost_local += cl;
                for (int off = 1; off < sB; off <<= 1) {               // all 32 lanes take part (sB divides 32: groups do not straddle warps)
#pragma unroll
                    for (int q = 0; q < 6; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
#pragma unroll
                    for (int q = 0; q < 3; ++q) g[q] += __shfl_xor_sync(0xffffffffu, g[q], off);
                    if (kFocal) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) Wf[q] += __shfl_xor_sync(0xffffffffu, Wf[q], off);
#pragma unroll
                        for (int q = 0; q < 4; ++q) f4[q] += __shfl_xor_sync(0xffffffffu, f4[q], off);
                    }
                }
                if (mine && sub == 0) {
                    // Marquardt damping D^2 = max(diag, 1e-6) / radius  (Ceres LM strategy with Jacobi scaling, min_lm_diagonal)
                    v[0] += fmax(v[0], 1e-6) * inv_radius;
                    v[3] += fmax(v[3], 1e-6) * inv_radius;
                    v[5] += fmax(v[5], 1e-6) * inv_radius;
                    sym3_inverse(v, Vinv);
                    gp[0] = g[0]; gp[1] = g[1]; gp[2] = g[2];
                    gpmax_local = fmax(gpmax_local, fmax(fabs(gp[0]), fmax(fabs(gp[1]), fabs(gp[2]))));
                    if (kFocal) {
                        // T = Wf V^-1 (2x3);  F -= T Wf^T,  rhs_f += T g_p - g_f   (the point is eliminated from the focal block too)
                        double* gw = P.pt_Wf + 6 * static_cast<size_t>(T.begin + u);
#pragma unroll
                        for (int q = 0; q < 6; ++q) gw[q] = Wf[q];
                        const double T0[3] = {Wf[0] * Vinv[0] + Wf[1] * Vinv[1] + Wf[2] * Vinv[2], Wf[0] * Vinv[1] + Wf[1] * Vinv[3] + Wf[2] * Vinv[4],
                                              Wf[0] * Vinv[2] + Wf[1] * Vinv[4] + Wf[2] * Vinv[5]};
                        const double T1[3] = {Wf[3] * Vinv[0] + Wf[4] * Vinv[1] + Wf[5] * Vinv[2], Wf[3] * Vinv[1] + Wf[4] * Vinv[3] + Wf[5] * Vinv[4],
                                              Wf[3] * Vinv[2] + Wf[4] * Vinv[4] + Wf[5] * Vinv[5]};
                        ff[0] += f4[0] - (T0[0] * Wf[0] + T0[1] * Wf[1] + T0[2] * Wf[2]);
                        ff[1] += -(T0[0] * Wf[3] + T0[1] * Wf[4] + T0[2] * Wf[5]);
                        ff[2] += f4[1] - (T1[0] * Wf[3] + T1[1] * Wf[4] + T1[2] * Wf[5]);
                        ff[3] += (T0[0] * gp[0] + T0[1] * gp[1] + T0[2] * gp[2]) - f4[2];
                        ff[4] += (T1[0] * gp[0] + T1[1] * gp[1] + T1[2] * gp[2]) - f4[3];
                        ff[5] += f4[2]; ff[6] += f4[3]; ff[7] += f4[0]; ff[8] += f4[1];
                    }
                }
            } else if (mine) {
                // a long track: the pre-pass reduced over the whole track
                const int d = P.items[T.begin + u].d;
                const double* rec = P.long_V + static_cast<size_t>(d - P.first_long) * (kFocal ? 15 : 9);
#pragma unroll
                for (int q = 0; q < 6; ++q) Vinv[q] = rec[q];
                gp[0] = rec[6]; gp[1] = rec[7]; gp[2] = rec[8];
                if (kFocal) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) Wf[q] = rec[9 + q];
                }
            }
            if (mine && sub == 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) ptV[6 * u + q] = static_cast<float>(Vinv[q]);
                ptg[3 * u] = gp[0]; ptg[3 * u + 1] = gp[1]; ptg[3 * u + 2] = gp[2];
                if (kFocal) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) ptWf[6 * u + q] = Wf[q];
                }
            }
            // bucket starts of the per-camera observation lists (the histogram of phase A, scanned in place)
            if (tid == kFusedThreads - 1) {
                int run = 0;
                for (int l = 0; l < T.w; ++l) { const int cnt = cam_start[l + 1]; cam_start[l] = run; cam_cursor[l] = run; run += cnt; }
                cam_start[T.w] = run;
            }
        }
        __syncthreads();
        // ---- C: this thread's observation again: Q = Jp V^-1, N = I - Q Jp^T, q = Q g_p - r; bucket by camera
        if (valid) {
            float* srow = stage + tid * kStageStride;
            const float* vi = ptV + 6 * unit;
            const float jp0 = srow[18], jp1 = srow[19], jp2 = srow[20], jp3 = srow[21], jp4 = srow[22], jp5 = srow[23];
            float Q[6];
            Q[0] = jp0 * vi[0] + jp1 * vi[1] + jp2 * vi[2]; Q[1] = jp0 * vi[1] + jp1 * vi[3] + jp2 * vi[4]; Q[2] = jp0 * vi[2] + jp1 * vi[4] + jp2 * vi[5];
            Q[3] = jp3 * vi[0] + jp4 * vi[1] + jp5 * vi[2]; Q[4] = jp3 * vi[1] + jp4 * vi[3] + jp5 * vi[4]; Q[5] = jp3 * vi[2] + jp4 * vi[4] + jp5 * vi[5];
            srow[12] = Q[0]; srow[13] = Q[1]; srow[14] = Q[2]; srow[15] = Q[3]; srow[16] = Q[4]; srow[17] = Q[5];
            if (f_mine >= 0 && diag_owner) {
                const double m00 = static_cast<double>(Q[0]) * jp0 + static_cast<double>(Q[1]) * jp1 + static_cast<double>(Q[2]) * jp2;
                const double m01 = static_cast<double>(Q[0]) * jp3 + static_cast<double>(Q[1]) * jp4 + static_cast<double>(Q[2]) * jp5;
                const double m11 = static_cast<double>(Q[3]) * jp3 + static_cast<double>(Q[4]) * jp4 + static_cast<double>(Q[5]) * jp5;
                srow[25] = static_cast<float>(1.0 - m00); srow[26] = static_cast<float>(-m01); srow[27] = static_cast<float>(1.0 - m11);
                const double* gp = ptg + 3 * unit;
                qbuf[2 * tid] = Q[0] * gp[0] + Q[1] * gp[1] + Q[2] * gp[2] - rbuf[2 * tid];
                qbuf[2 * tid + 1] = Q[3] * gp[0] + Q[4] * gp[1] + Q[5] * gp[2] - rbuf[2 * tid + 1];
                cam_obs[atomicAdd(cam_cursor + lcam, 1)] = static_cast<uint16_t>(tid);
            }
        }
        __syncthreads();
        // the next tile's camera / free-index / slot tables, copied asynchronously under phase E
        if (ti_next < P.n_tiles) {
            TileHdr& N = hdr[cur ^ 1];
            const int wn = N.T.w, cbn = N.T.cam_begin, sbn = N.T.slot_begin;
            if (tid < wn) cp_async4(N.cams + tid, P.tile_cams + cbn + tid);
            else if (tid >= 32 && tid - 32 < wn) cp_async4(N.free_ + (tid - 32), P.tile_free + cbn + (tid - 32));
            for (int i = tid; i < wn * (wn + 1) / 2; i += kFusedThreads) cp_async4(N.slots + i, P.tile_slots + sbn + i);
        }
        // ---- E: work queue: camera items first, then one item per unit
        {
            constexpr int kCamSplit = 4;                    // every camera part is summed by four items (a quarter of the list each)
            constexpr int kCamParts = (kFocal ? 6 : 4) * kCamSplit;   // 3 row pairs of the diagonal block | rhs, g_c, diag U (| 2 border columns)
            const int n_items = kCamParts + T.n_runs;
            for (;;) {
                int item = 0;
                if (lane == 0) item = atomicAdd(misc + 1, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= n_items) break;
                if (item < kCamParts) {
                    // lane = local camera, item = (part, quarter of the camera's observation list); the four quarters meet in
                    // the shared accumulators (at most four adders per address)
                    const int l = lane, part = item / kCamSplit, quarter = item - part * kCamSplit;
                    if (l < T.w) {
                        const int pb = cam_start[l] + quarter, pe = cam_start[l + 1];
                        if (part < 3) {
                            float* blk = acc + (l * (l + 1) / 2 + l) * kBlkStride;
                            if (part == 0) diag_row_pair<0>(stage, cam_obs, pb, pe, kCamSplit, blk);
                            else if (part == 1) diag_row_pair<1>(stage, cam_obs, pb, pe, kCamSplit, blk);
                            else diag_row_pair<2>(stage, cam_obs, pb, pe, kCamSplit, blk);
                        } else if (part == 3) {
                            // rhs = sum Jc^T q, g_c = sum Jc^T r (fp64), diag U = sum Jc^2 (only scales the damping: fp32 sums)
                            double ar[6] = {0, 0, 0, 0, 0, 0}, ag[6] = {0, 0, 0, 0, 0, 0};
                            float au[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            for (int pp = pb; pp < pe; pp += kCamSplit) {
                                const int row = cam_obs[pp];
                                const float* sr = stage + row * kStageStride;
                                const double q0 = qbuf[2 * row], q1 = qbuf[2 * row + 1], e0 = rbuf[2 * row], e1 = rbuf[2 * row + 1];
#pragma unroll
                                for (int i = 0; i < 6; ++i) {
                                    const float f0 = sr[i], f1 = sr[6 + i];
                                    const double c0 = f0, c1 = f1;
                                    ar[i] += c0 * q0 + c1 * q1;
                                    ag[i] += c0 * e0 + c1 * e1;
                                    au[i] += f0 * f0 + f1 * f1;
                                }
                            }
                            if (pb < pe) {
                                double ud[6];
#pragma unroll
                                for (int i = 0; i < 6; ++i) ud[i] = au[i];
                                smem_add6(camacc + l * CV, ar);
                                smem_add6(camacc + l * CV + 6, ag);
                                smem_add6(camacc + l * CV + 12, ud);
                            }
                        } else if (kFocal) {
                            // border column (part - 4) of B_c = sum Jc^T (Jf - Q Wf^T),  Jf = diag(xp, yp)
                            double a[6] = {0, 0, 0, 0, 0, 0};
                            const int col = part - 4;
                            for (int pp = pb; pp < pe; pp += kCamSplit) {
                                const int row = cam_obs[pp];
                                const float* sr = stage + row * kStageStride;
                                const int un = split ? row >> 5 : static_cast<int>(__ldg(P.obs_lpt + T.obs_begin + row));
                                const double* Wf = ptWf + 6 * un + 3 * col;
                                const double g0 = sr[12] * Wf[0] + sr[13] * Wf[1] + sr[14] * Wf[2];      // (Q Wf^T)[0][col]
                                const double g1 = sr[15] * Wf[0] + sr[16] * Wf[1] + sr[17] * Wf[2];      // (Q Wf^T)[1][col]
                                const double e0 = (col == 0 ? xybuf[2 * row] : 0.0) - g0, e1 = (col == 1 ? xybuf[2 * row + 1] : 0.0) - g1;
#pragma unroll
                                for (int i = 0; i < 6; ++i) a[i] += static_cast<double>(sr[i]) * e0 + static_cast<double>(sr[6 + i]) * e1;
                            }
                            if (pb < pe) smem_add6(camacc + l * CV + 18 + 6 * col, a);
                        }
                    }
                    continue;
                }
                // ---- run item: lanes = camera pairs (x < y) of a run of points with identical camera lists; the products of
                //      the whole run are summed in registers (packed fp32x2 FMAs) and added to the shared block once
                const int ri = item - kCamParts;
                const uint32_t rn = ri < kTileRunItems ? run_items[ri] : __ldg(P.runs + T.run_begin + ri);
                const int u0 = static_cast<int>(rn & 0xFFFFu), nrun = static_cast<int>((rn >> 16) & 0xFFu);
                const uint32_t ui = unit_info[u0];
                const int nA = static_cast<int>((ui >> 16) & 0xFFu), nB = static_cast<int>(ui >> 24);
                const int npairs = nB > 0 ? nA * nB : nA * (nA - 1) / 2;
                {
                    const int ql = static_cast<int>(rn >> 24) * 32 + lane;
                    if (ql >= npairs) continue;
                    int x, y;
                    if (nB > 0) {
                        x = ql / nB; y = nA + (ql - x * nB);
                    } else {
                        y = static_cast<int>((1.0f + sqrtf(1.0f + 8.0f * static_cast<float>(ql))) * 0.5f);
                        y -= (y * (y - 1) / 2 > ql) ? 1 : 0;
                        y += ((y + 1) * y / 2 <= ql) ? 1 : 0;
                        x = ql - y * (y - 1) / 2;
                    }
                    float2 av[18];
#pragma unroll
                    for (int e = 0; e < 18; ++e) av[e] = make_float2(0.f, 0.f);
                    int lx = 0, ly = 0;
                    for (int j = 0; j < nrun; ++j) {
                        const int ob = static_cast<int>(unit_info[u0 + j] & 0xFFFFu);
                        const float4* sx = reinterpret_cast<const float4*>(stage + (ob + x) * kStageStride);
                        const float4* sy = reinterpret_cast<const float4*>(stage + (ob + y) * kStageStride);
                        if (j == 0) { lx = __float_as_int(sx[6].x); ly = __float_as_int(sy[6].x); }
                        if (lx < 0 || ly < 0) break;                            // a constant camera: the same for the whole run
                        const float4 xq0 = sx[3], xq1 = sx[4];                  // Q_x = (xq0.xyzw, xq1.xy)
                        const float4 yp0 = sy[4], yp1 = sy[5];                  // Jp_y = (yp0.zw, yp1.xyzw)
                        const float m00 = -(xq0.x * yp0.z + xq0.y * yp0.w + xq0.z * yp1.x);
                        const float m01 = -(xq0.x * yp1.y + xq0.y * yp1.z + xq0.z * yp1.w);
                        const float m10 = -(xq0.w * yp0.z + xq1.x * yp0.w + xq1.y * yp1.x);
                        const float m11 = -(xq0.w * yp1.y + xq1.x * yp1.z + xq1.y * yp1.w);
                        const float4 a0 = sy[0], a1 = sy[1], a2 = sy[2];        // Jc_y rows: (a0.xyzw a1.xy) | (a1.zw a2.xyzw)
                        const float2 d00 = make_float2(m00, m00), d01 = make_float2(m01, m01), d10 = make_float2(m10, m10), d11 = make_float2(m11, m11);
                        float2 T0[3], T1[3];
                        T0[0] = __ffma2_rn(d01, make_float2(a1.z, a1.w), __fmul2_rn(d00, make_float2(a0.x, a0.y)));
                        T0[1] = __ffma2_rn(d01, make_float2(a2.x, a2.y), __fmul2_rn(d00, make_float2(a0.z, a0.w)));
                        T0[2] = __ffma2_rn(d01, make_float2(a2.z, a2.w), __fmul2_rn(d00, make_float2(a1.x, a1.y)));
                        T1[0] = __ffma2_rn(d11, make_float2(a1.z, a1.w), __fmul2_rn(d10, make_float2(a0.x, a0.y)));
                        T1[1] = __ffma2_rn(d11, make_float2(a2.x, a2.y), __fmul2_rn(d10, make_float2(a0.z, a0.w)));
                        T1[2] = __ffma2_rn(d11, make_float2(a2.z, a2.w), __fmul2_rn(d10, make_float2(a1.x, a1.y)));
                        const float4 b0 = sx[0], b1 = sx[1], b2 = sx[2];        // Jc_x
                        const float jx0[6] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y}, jx1[6] = {b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const float2 c0 = make_float2(jx0[i], jx0[i]), c1 = make_float2(jx1[i], jx1[i]);
#pragma unroll
                            for (int c = 0; c < 3; ++c) av[3 * i + c] = __ffma2_rn(c1, T1[c], __ffma2_rn(c0, T0[c], av[3 * i + c]));
                        }
                    }
                    if (lx < 0 || ly < 0) continue;
                    smem_add_block(acc + (ly * (ly + 1) / 2 + lx) * kBlkStride, av);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- flush the tile: 6x6 blocks with vector reductions, camera vectors with fp64 reductions
        for (int u = tid; u < nb * 9; u += kFusedThreads) {
            const int b = u / 9, q4 = u - 9 * b;
            const int slot = H.slots[b];
            if (slot < 0) continue;
            const float4 v4 = reinterpret_cast<const float4*>(acc + b * kBlkStride)[q4];
            atomicAdd(reinterpret_cast<float4*>(P.sblk + static_cast<size_t>(slot) * 36) + q4, v4);
        }
        for (int u = tid; u < T.w * CV; u += kFusedThreads) {
            const int l = u / CV, e = u - l * CV;
            const int f = H.free_[l];
            if (f < 0) continue;
            double* dst;
            if (e < 6) dst = tail + P.tl.rhs + f * 6 + e;
            else if (e < 12) dst = tail + P.tl.gc + f * 6 + (e - 6);
            else if (e < 18) dst = tail + P.tl.udiag + f * 6 + (e - 12);
            else dst = tail + (e < 24 ? P.tl.B0 : P.tl.B1) + f * 6 + (e < 24 ? e - 18 : e - 24);
            atomicAdd(dst, camacc[u]);
        }
        if (tid <= kTileCams) cam_start[tid] = 0;               // the histogram of the next tile
        __syncthreads();
        cur ^= 1;
    }
    // ---- scalars: one fp64 atomic per CTA
    __shared__ double sh_c[kFusedThreads / 32], sh_g[kFusedThreads / 32];
    cost_local = warp_sum(cost_local);
    gpmax_local = warp_max(gpmax_local);
    if (lane == 0) { sh_c[warp] = cost_local; sh_g[warp] = gpmax_local; }
    __syncthreads();
    if (kFocal) {
        __shared__ double sh_f[kFusedThreads / 32][9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double v = warp_sum(ff[k]);
            if (lane == 0) sh_f[warp][k] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
            for (int k = 0; k < kFusedThreads / 32; ++k) v += sh_f[k][tid];
            atomicAdd(tail + P.tl.ff + tid, v);
        }
    }
    if (tid == 0) {
        double c = 0.0, g = 0.0;
        for (int k = 0; k < kFusedThreads / 32; ++k) { c += sh_c[k]; g = fmax(g, sh_g[k]); }
        atomicAdd(tail + P.tl.scal, c);
        // max of non-negative doubles == max of their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(tail + P.tl.gpm + P.gpm_slot), static_cast<unsigned long long>(__double_as_longlong(g)));
    }
}

// Dense fp64 copy of the block-sparse S for the dense Cholesky: upper block triangle, row-major (= the column-major lower
// triangle cuSOLVER reads), Marquardt damping max(diag U, 1e-6) / radius added on the diagonal.  S must be zeroed.
__global__ void expand_dense_kernel(Problem P, double inv_radius, double* __restrict__ S) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.n_blocks * 36) return;
    const int b = idx / 36, e = idx - 36 * b, i = e / 6, j = e - 6 * i;
    const int fa = __ldg(P.blk_row + b), fb = __ldg(P.blk_col + b);
    const size_t n6 = static_cast<size_t>(P.n_free) * 6;
    double v = static_cast<double>(P.sblk[idx]);
    if (fa == fb) {
        v = static_cast<double>(P.sblk[b * 36 + (i <= j ? 6 * i + j : 6 * j + i)]);     // a diagonal block holds its upper triangle
        if (i == j) v += fmax(P.tail[P.tl.udiag + fa * 6 + i], 1e-6) * inv_radius;
    }
    S[(static_cast<size_t>(fa) * 6 + i) * n6 + static_cast<size_t>(fb) * 6 + j] = v;
}

// ------------------------------------------------------------------------------------------------ back-substitution
// dp = -V^-1 (g_p + sum_obs W^T dc);  candidate point = point + dp;  accumulates
// out[0] += model decrease  -(r.Jd + 1/2 |Jd|^2),  out[1] += |dp|^2,  out[2] += |point|^2
__global__ void __launch_bounds__(256)
backsub_kernel(Problem P, int first_point, double inv_radius, const double* __restrict__ dc /*[n_free*6]*/, double* __restrict__ pts_new,
               double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    double acc_model = 0.0, acc_dp = 0.0, acc_x = 0.0;
    // warp index and track bounds go through a shuffle so that the compiler knows they are uniform across the warp: otherwise every
    // shuffle of the reductions below is compiled as a convergence barrier (WARPSYNC.COLLECTIVE), one at a time
    const int wid = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    for (int d = first_point + blockIdx.x * wpb + wid; d < P.n_pts; d += gridDim.x * wpb) {
        const int p = P.pt_order[d];
        const int beg = __shfl_sync(0xffffffffu, P.pt_start[d], 0), end = __shfl_sync(0xffffffffu, P.pt_start[d + 1], 0);
        const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
        if (beg == end) {
            if (lane == 0) { pts_new[3 * p] = X[0]; pts_new[3 * p + 1] = X[1]; pts_new[3 * p + 2] = X[2]; }
            continue;
        }
        double Vinv[6], gp[3], unused = 0.0;
        LaneObs A;
        point_pass1(P, beg, end, lane, X, inv_radius, Vinv, gp, A, unused);
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
                    const double dd = dc[A.f * 6 + i];
                    jd0 += A.Jc[i] * dd; jd1 += A.Jc[6 + i] * dd;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] += A.Jp[k] * jd0 + A.Jp[3 + k] * jd1;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = gp[k] + warp_sum(t[k]);
        if (focal) {
            const double* wf = P.pt_Wf + 6 * static_cast<size_t>(d);
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
                        const double dd = dc[A.f * 6 + i];
                        jd0 += A.Jc[i] * dd; jd1 += A.Jc[6 + i] * dd;
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

// The same back-substitution for the NORMAL tiles (points with at most 32 views), in the tile layout of the linearisation:
// thread = observation for the geometry (every lane busy, instead of one warp per point with a third of its lanes), thread =
// point for the 3x3 work.  Long tracks and unobserved points stay with backsub_kernel (device points >= first_long).
__global__ void __launch_bounds__(kTileObs, 1)
backsub_tile_kernel(Problem P, double inv_radius, const double* __restrict__ dc, double* __restrict__ pts_new, double* __restrict__ out) {
    __shared__ float sJp[kTileObs][6];
    __shared__ double sR[kTileObs][2], sJd[kTileObs][2], sXy[kTileObs][2];
    __shared__ double sh[3][kTileObs / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool focal = P.refine_focal != 0;
    const double df0 = focal ? dc[static_cast<size_t>(P.n_free) * 6] : 0.0, df1 = focal ? dc[static_cast<size_t>(P.n_free) * 6 + 1] : 0.0;
    double acc_model = 0.0, acc_dp = 0.0, acc_x = 0.0;
    for (int ti = blockIdx.x; ti < P.n_tiles; ti += gridDim.x) {
        const Tile T = P.tiles[ti];
        if (T.flags & kTileSplit) continue;
        __syncthreads();
        if (tid < T.n_obs) {
            const int o = T.obs_begin + tid;
            const int cam = __ldg(P.obs_cam + o);
            const int f = __ldg(P.cam_free + cam);
            const int p = __ldg(P.pt_order + T.begin + static_cast<int>(__ldg(P.obs_lpt + o)));
            const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
            const double2 uv = __ldg(reinterpret_cast<const double2*>(P.obs_uv) + o);
            double r[2], Jc[12], Jp[6], xy[2];
            obs_eval<true>(P.pre[cam], X, uv.x, uv.y, P.fx, P.fy, r, Jc, Jp, xy);
            double jd0 = 0.0, jd1 = 0.0;
            if (f >= 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double dd = dc[f * 6 + i];
                    jd0 += static_cast<double>(static_cast<float>(Jc[i])) * dd;        // the fp32 Jacobians of the linearisation
                    jd1 += static_cast<double>(static_cast<float>(Jc[6 + i])) * dd;
                }
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) sJp[tid][k] = static_cast<float>(Jp[k]);
            sR[tid][0] = r[0]; sR[tid][1] = r[1];
            sJd[tid][0] = jd0; sJd[tid][1] = jd1;
            sXy[tid][0] = xy[0]; sXy[tid][1] = xy[1];
        }
        __syncthreads();
        if (tid < T.end - T.begin) {
            const int d = T.begin + tid;
            const int p = __ldg(P.pt_order + d);
            const int ob = __ldg(P.pt_start + d) - T.obs_begin, k = __ldg(P.pt_start + d + 1) - __ldg(P.pt_start + d);
            const double X[3] = {P.pts[3 * p], P.pts[3 * p + 1], P.pts[3 * p + 2]};
            double v[6] = {0, 0, 0, 0, 0, 0}, t[3] = {0, 0, 0};
            for (int j = 0; j < k; ++j) {
                const double j0 = sJp[ob + j][0], j1 = sJp[ob + j][1], j2 = sJp[ob + j][2], j3 = sJp[ob + j][3], j4 = sJp[ob + j][4], j5 = sJp[ob + j][5];
                const double a0 = sR[ob + j][0] + sJd[ob + j][0], a1 = sR[ob + j][1] + sJd[ob + j][1];     // g_p + Jp^T (Jc dc)
                v[0] += j0 * j0 + j3 * j3; v[1] += j0 * j1 + j3 * j4; v[2] += j0 * j2 + j3 * j5;
                v[3] += j1 * j1 + j4 * j4; v[4] += j1 * j2 + j4 * j5; v[5] += j2 * j2 + j5 * j5;
                t[0] += j0 * a0 + j3 * a1; t[1] += j1 * a0 + j4 * a1; t[2] += j2 * a0 + j5 * a1;
            }
            v[0] += fmax(v[0], 1e-6) * inv_radius;
            v[3] += fmax(v[3], 1e-6) * inv_radius;
            v[5] += fmax(v[5], 1e-6) * inv_radius;
            double Vinv[6];
            sym3_inverse(v, Vinv);
            if (focal) {
                const double* wf = P.pt_Wf + 6 * static_cast<size_t>(d);
#pragma unroll
                for (int q = 0; q < 3; ++q) t[q] += wf[q] * df0 + wf[3 + q] * df1;
            }
            double dp[3];
            dp[0] = -(Vinv[0] * t[0] + Vinv[1] * t[1] + Vinv[2] * t[2]);
            dp[1] = -(Vinv[1] * t[0] + Vinv[3] * t[1] + Vinv[4] * t[2]);
            dp[2] = -(Vinv[2] * t[0] + Vinv[4] * t[1] + Vinv[5] * t[2]);
            pts_new[3 * p] = X[0] + dp[0]; pts_new[3 * p + 1] = X[1] + dp[1]; pts_new[3 * p + 2] = X[2] + dp[2];
            acc_dp += dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
            acc_x += X[0] * X[0] + X[1] * X[1] + X[2] * X[2];
            for (int j = 0; j < k; ++j) {
                const double jd0 = sJd[ob + j][0] + sJp[ob + j][0] * dp[0] + sJp[ob + j][1] * dp[1] + sJp[ob + j][2] * dp[2] + sXy[ob + j][0] * df0;
                const double jd1 = sJd[ob + j][1] + sJp[ob + j][3] * dp[0] + sJp[ob + j][4] * dp[1] + sJp[ob + j][5] * dp[2] + sXy[ob + j][1] * df1;
                acc_model -= sR[ob + j][0] * jd0 + sR[ob + j][1] * jd1 + 0.5 * (jd0 * jd0 + jd1 * jd1);
            }
        }
    }
    acc_model = warp_sum(acc_model); acc_dp = warp_sum(acc_dp); acc_x = warp_sum(acc_x);
    __syncthreads();
    if (lane == 0) { sh[0][warp] = acc_model; sh[1][warp] = acc_dp; sh[2][warp] = acc_x; }
    __syncthreads();
    if (tid < 3) {
        double s2 = 0.0;
        for (int k = 0; k < kTileObs / 32; ++k) s2 += sh[tid][k];
        atomicAdd(&out[tid], s2);
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

// The 72-byte record of one LM iteration the host reads back (one D2H per iteration instead of the tail / the step / the
// camera mirror): rec[0] cost at the linearisation point, [1] max |gradient| (cameras, points of every rank, focal block),
// [2] |dc|^2, [3] |x_c|^2 over the free cameras, [4] model decrease, [5] |dp|^2, [6] |x_p|^2, [7] cost of the candidate,
// [8] status of the linear solve (0 = positive definite).
__global__ void __launch_bounds__(256)
lm_record_kernel(Problem P, const double* __restrict__ cams, const double* __restrict__ dc, const double* __restrict__ small,
                 const int* __restrict__ info, int n_info, int n_ranks, double* __restrict__ rec) {
    const int n6 = P.n_free * 6;
    double gmax = 0.0, dc2 = 0.0, xc2 = 0.0, bad = 0.0;
    for (int i = threadIdx.x; i < n6; i += blockDim.x) {
        gmax = fmax(gmax, fabs(P.tail[P.tl.gc + i]));
        dc2 += dc[i] * dc[i];
    }
    for (int i = threadIdx.x; i < P.n_cams * 6; i += blockDim.x)
        if (P.cam_free[i / 6] >= 0) xc2 += cams[i] * cams[i];
    for (int i = threadIdx.x; i < n_ranks; i += blockDim.x) gmax = fmax(gmax, P.tail[P.tl.gpm + i]);
    for (int i = threadIdx.x; i < n_info; i += blockDim.x) bad = fmax(bad, fabs(static_cast<double>(info[i])));
    if (P.refine_focal && threadIdx.x == 0) gmax = fmax(gmax, fmax(fabs(P.tail[P.tl.ff + 5]), fabs(P.tail[P.tl.ff + 6])));
    __shared__ double sh[4][8];
    gmax = warp_max(gmax); bad = warp_max(bad); dc2 = warp_sum(dc2); xc2 = warp_sum(xc2);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = gmax; sh[1][threadIdx.x >> 5] = dc2; sh[2][threadIdx.x >> 5] = xc2; sh[3][threadIdx.x >> 5] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0, a = 0.0, b = 0.0, e = 0.0;
        for (int k = 0; k < (blockDim.x >> 5); ++k) { g = fmax(g, sh[0][k]); a += sh[1][k]; b += sh[2][k]; e = fmax(e, sh[3][k]); }
        rec[0] = P.tail[P.tl.scal]; rec[1] = g; rec[2] = a; rec[3] = b;
        rec[4] = small[0]; rec[5] = small[1]; rec[6] = small[2]; rec[7] = small[3]; rec[8] = e;
    }
}

__global__ void copy_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}


// Measurements (and camera indices) from the caller's observation order into device order: out[i] = in[obs_orig[i]].
// msfm_ba_create / msfm_ba_update upload the caller's arrays as they are; the gather runs here at HBM speed instead of on
// one host thread.
__global__ void permute_obs_kernel(int n_obs, const int32_t* __restrict__ obs_orig, const double2* __restrict__ uv_in,
                                   const int32_t* __restrict__ cam_in, double2* __restrict__ uv_out, int32_t* __restrict__ cam_out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_obs; i += gridDim.x * blockDim.x) {
        const int o = __ldg(obs_orig + i);
        uv_out[i] = __ldg(uv_in + o);
        if (cam_in) cam_out[i] = __ldg(cam_in + o);
    }
}
// obs_pt in device order (the caller's point index of every device observation) from the point tables
__global__ void fill_obs_pt_kernel(int n_pts, const int32_t* __restrict__ pt_start, const int32_t* __restrict__ pt_order,
                                   int32_t* __restrict__ obs_pt) {
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n_pts; d += gridDim.x * blockDim.x) {
        const int p = __ldg(pt_order + d);
        for (int a = __ldg(pt_start + d); a < __ldg(pt_start + d + 1); ++a) obs_pt[a] = p;
    }
}
}  // namespace ba

// ------------------------------------------------------------------------------------------------ launchers
using namespace ba;
cudaError_t ba_launch_cam_prep(const double* cams, int n_cams, CamPre* pre, cudaStream_t st) {
    if (n_cams <= 0) return cudaSuccess;
    { cam_prep_kernel<<<(n_cams + 127) / 128, 128, 0, st>>>(cams, n_cams, pre); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_evaluate(const Problem& P, double* r_out, float* J_out, double* cost, int num_sms, cudaStream_t st) {
    if (P.n_obs <= 0) return cudaSuccess;
    int grid = (P.n_obs + 255) / 256;
    if (grid > num_sms * 8) grid = num_sms * 8;
    { evaluate_kernel<<<grid, 256, 0, st>>>(P, r_out, J_out, cost); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_track_errors(const Problem& P, double* err, int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    int grid = (P.n_pts + 7) / 8;
    if (grid > num_sms * 8) grid = num_sms * 8;
    { track_error_kernel<<<grid, 256, 0, st>>>(P, err); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_filter_stats(const Problem& P, double max_err, uint8_t* keep, double* err, int32_t* kept, double* angle, int num_sms,
                                   cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    int grid = (P.n_pts + 7) / 8;
    if (grid > num_sms * 8) grid = num_sms * 8;
    { filter_stats_kernel<<<grid, 256, 0, st>>>(P, max_err, keep, err, kept, angle); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
size_t ba_fused_smem_bytes(bool focal) { return fused_smem_layout(focal).total; }
// The system buffers (tail, tile counter, sblk) must be zero.
cudaError_t ba_launch_linearize(const Problem& P, double inv_radius, int num_sms, cudaStream_t st) {
    if (P.n_tiles <= 0) return cudaSuccess;
    if (P.n_long > 0) {
        int g = (P.n_long + 7) / 8;
        if (g > num_sms * 4) g = num_sms * 4;
        if (P.refine_focal) { long_track_prepass_kernel<true><<<g, 256, 0, st>>>(P, inv_radius); MSFM_COUNT_LAUNCH(); }
        else { long_track_prepass_kernel<false><<<g, 256, 0, st>>>(P, inv_radius); MSFM_COUNT_LAUNCH(); }
    }
    const size_t smem = fused_smem_layout(P.refine_focal != 0).total;
    cudaError_t e;
    if (P.refine_focal) e = cudaFuncSetAttribute(fused_linearize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    else e = cudaFuncSetAttribute(fused_linearize_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (kCtasPerSm > 1) {      // two CTAs per SM need (nearly) the whole shared memory of the SM
        if (P.refine_focal) e = cudaFuncSetAttribute(fused_linearize_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        else e = cudaFuncSetAttribute(fused_linearize_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
    }
    const int grid = P.n_tiles < num_sms * kCtasPerSm ? P.n_tiles : num_sms * kCtasPerSm;
    if (P.refine_focal) { fused_linearize_kernel<true><<<grid, kFusedThreads, smem, st>>>(P, inv_radius); MSFM_COUNT_LAUNCH(); }
    else { fused_linearize_kernel<false><<<grid, kFusedThreads, smem, st>>>(P, inv_radius); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_expand_dense(const Problem& P, double inv_radius, double* S, cudaStream_t st) {
    if (P.n_blocks <= 0) return cudaSuccess;
    const int n = P.n_blocks * 36;
    { expand_dense_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, inv_radius, S); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_backsub(const Problem& P, double inv_radius, const double* dc, double* pts_new, double* out,
                              int num_sms, cudaStream_t st) {
    if (P.n_pts <= 0) return cudaSuccess;
    // normal tiles: thread-per-observation kernel; long tracks + unobserved points (device points >= first_long): warp per point
    if (P.first_long > 0 && P.n_tiles > 0) {
        const int grid = P.n_tiles < num_sms * kCtasPerSm ? P.n_tiles : num_sms * kCtasPerSm;
        { backsub_tile_kernel<<<grid, kTileObs, 0, st>>>(P, inv_radius, dc, pts_new, out); MSFM_COUNT_LAUNCH(); }
    }
    const int rest = P.n_pts - P.first_long;
    if (rest > 0) {
        int grid = (rest + 7) / 8;
        if (grid > num_sms * 8) grid = num_sms * 8;
        { backsub_kernel<<<grid, 256, 0, st>>>(P, P.first_long, inv_radius, dc, pts_new, out); MSFM_COUNT_LAUNCH(); }
    }
    return cudaGetLastError();
}
cudaError_t ba_launch_update_cams(const double* cams, const int32_t* cam_free, int n_cams, const double* dc,
                                  double* cams_new, cudaStream_t st) {
    if (n_cams <= 0) return cudaSuccess;
    { update_cams_kernel<<<(n_cams * 6 + 127) / 128, 128, 0, st>>>(cams, cam_free, n_cams, dc, cams_new); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_lm_record(const Problem& P, const double* cams, const double* dc, const double* small, const int* info, int n_info,
                                int n_ranks, double* rec, cudaStream_t st) {
    { lm_record_kernel<<<1, 256, 0, st>>>(P, cams, dc, small, info, n_info, n_ranks, rec); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_copy(const double* src, double* dst, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    { copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

cudaError_t ba_launch_permute_obs(int n_obs, const int32_t* obs_orig, const double* uv_in, const int32_t* cam_in, double* uv_out,
                                  int32_t* cam_out, cudaStream_t st) {
    if (n_obs <= 0) return cudaSuccess;
    const int grid = std::min((n_obs + 255) / 256, 148 * 16);
    { permute_obs_kernel<<<grid, 256, 0, st>>>(n_obs, obs_orig, reinterpret_cast<const double2*>(uv_in), cam_in,
                                             reinterpret_cast<double2*>(uv_out), cam_out); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t ba_launch_fill_obs_pt(int n_pts, const int32_t* pt_start, const int32_t* pt_order, int32_t* obs_pt, cudaStream_t st) {
    if (n_pts <= 0) return cudaSuccess;
    const int grid = std::min((n_pts + 255) / 256, 148 * 16);
    { fill_obs_pt_kernel<<<grid, 256, 0, st>>>(n_pts, pt_start, pt_order, obs_pt); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

}  // namespace msfm
