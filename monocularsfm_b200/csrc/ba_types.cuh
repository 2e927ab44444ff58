// Device-side data layout of the B-path (bundle adjustment).  See DESIGN.md §"B-path layout".
#pragma once
#include <cstdint>

namespace msfm {
namespace ba {

// Per-camera quantities recomputed once per parameter update (fp64): rotation matrix of Ceres'
// AngleAxisRotatePoint and the four scalar coefficients of its derivative.
struct CamPre {
    double R[9];
    double w[3];
    double t[3];
    double a, b, a1, b1;
    double pad;
};

// SoA view of a BundleData (include/Optimizer/BundleData.h:19-65) flattened for the device.
struct Problem {
    int32_t n_cams, n_pts, n_obs, n_free;
    double fx, fy;
    const CamPre* pre;          // [n_cams]
    const double* pts;          // [n_pts][3]
    const double* obs_uv;       // [n_obs][2], centred by (cx, cy)
    const int32_t* obs_cam;     // [n_obs]
    const int32_t* obs_pt;      // [n_obs], non-decreasing
    const int32_t* pt_start;    // [n_pts+1] CSR over observations
    const int32_t* cam_free;    // [n_cams] index among the free cameras or -1 (constant pose)
    unsigned long long* gpmax_bits;   // max |g_p| as the bit pattern of a non-negative double
    // ---- gather structures (built once per problem: the sparsity pattern does not change between LM iterations)
    const int32_t* cam_obs_start;   // [n_free+1] CSR: observations of every free camera
    const int32_t* cam_obs_list;    // [#observations of free cameras]
    const int32_t* blk_start;       // [n_free*n_free+1] CSR over camera-pair blocks (fa < fb): co-observing tuples
    const int2* blk_tuples;         // (obs_a, obs_b) with cam_free[obs_cam[obs_a]] = fa < fb = cam_free[obs_cam[obs_b]]
    const int32_t* blk_list;        // [n_blk_list] indices fa*n_free+fb of the blocks with at least one tuple
    int32_t n_blk_list;
    // ---- per-linearisation intermediates
    float* obs_J;                   // [n_obs][18]  Jc (2x6) | Jp (2x3), fp32
    double* obs_r;                  // [n_obs][2]
    double* pt_Vinv;                // [n_pts][6]   inverse of the damped point block (symmetric)
    double* pt_gp;                  // [n_pts][3]
    // ---- shared focal block (BundleAutoDiffCostFunction, CeresBundleOptimizer.cpp:76-121; refine_focal_length)
    int32_t refine_focal;           // 0: (fx, fy) constant; 1: one shared 2-parameter block, bordering the camera system
    double* pt_Wf;                  // [n_pts][6]   Wf = sum_obs Jf^T Jp (2x3), Jf = d r / d (fx, fy) = diag(xp, yp)
};

}  // namespace ba
}  // namespace msfm
