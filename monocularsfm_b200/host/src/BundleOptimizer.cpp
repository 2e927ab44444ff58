// BundleData::Debug and CeresBundelOptimizer::Optimize on top of the C-ABI B-path.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "DeviceContext.h"
#include "Optimizer/BundleData.h"
#include "Optimizer/CeresBundleOptimizer.h"

using namespace MonocularSfM;

void initLogging() {}

namespace {
// AoS hash maps -> SoA arrays.  Landmarks are visited in id order so the flattening is deterministic
// (the reference iterates in hash order, CeresBundleOptimizer.cpp:213; the optimum does not depend on it).
struct Flat {
    std::vector<image_t> cam_ids;
    std::vector<point3D_t> pt_ids;
    std::vector<double> cams, pts, uv;
    std::vector<int32_t> obs_cam, obs_pt;
    std::vector<uint8_t> cam_const;
    double fx, fy, cx, cy;
};

// A pose vector as the reference hands it to Ceres: `rvec.data` taken as double[3] (CeresBundleOptimizer.cpp:230-233), which
// holds for a 3x1 and for a 1x3 Mat alike.  Anything else is a caller error, reported instead of read out of bounds.
double* PoseData(cv::Mat& v, const char* what) {
    if (v.type() != CV_64F || v.rows * v.cols != 3 || !v.data)
        throw std::runtime_error(std::string("CeresBundelOptimizer: ") + what + " must be a CV_64F Mat with 3 elements");
    return reinterpret_cast<double*>(v.data);
}

Flat Flatten(BundleData& bd) {
    assert(bd.K.type() == CV_64F);                                           // :190
    Flat f;
    f.fx = bd.K.at<double>(0, 0); f.fy = bd.K.at<double>(1, 1);              // :197-200
    f.cx = bd.K.at<double>(0, 2); f.cy = bd.K.at<double>(1, 2);
    for (auto& el : bd.camera_poses) f.cam_ids.push_back(el.first);
    std::sort(f.cam_ids.begin(), f.cam_ids.end());
    std::unordered_map<image_t, int> cam_index;
    for (size_t i = 0; i < f.cam_ids.size(); ++i) cam_index[f.cam_ids[i]] = static_cast<int>(i);
    f.cams.resize(f.cam_ids.size() * 6);
    f.cam_const.assign(f.cam_ids.size(), 0);
    for (size_t i = 0; i < f.cam_ids.size(); ++i) {
        BundleData::CameraPose& cp = bd.camera_poses[f.cam_ids[i]];
        const double* rv = PoseData(cp.rvec, "rvec");
        const double* tv = PoseData(cp.tvec, "tvec");
        for (int k = 0; k < 3; ++k) {
            f.cams[6 * i + k] = rv[k];
            f.cams[6 * i + 3 + k] = tv[k];
        }
        if (bd.constant_camera_pose.count(f.cam_ids[i])) f.cam_const[i] = 1;   // :256-260
    }
    for (auto& el : bd.landmarks) f.pt_ids.push_back(el.first);
    std::sort(f.pt_ids.begin(), f.pt_ids.end());
    f.pts.resize(f.pt_ids.size() * 3);
    for (size_t p = 0; p < f.pt_ids.size(); ++p) {
        BundleData::Landmark& lm = bd.landmarks[f.pt_ids[p]];
        for (int k = 0; k < 3; ++k) f.pts[3 * p + k] = lm.point3D(k);
        for (const BundleData::Measurement& m : lm.measurements) {
            f.obs_cam.push_back(cam_index.at(m.image_id));
            f.obs_pt.push_back(static_cast<int32_t>(p));
            f.uv.push_back(m.point2D(0) - f.cx);                               // :221-222
            f.uv.push_back(m.point2D(1) - f.cy);
        }
    }
    return f;
}

msfm_ba_problem Describe(const Flat& f, bool refine_focal_length) {
    msfm_ba_problem pr;
    pr.n_cams = static_cast<int32_t>(f.cam_ids.size());
    pr.n_pts = static_cast<int32_t>(f.pt_ids.size());
    pr.n_obs = static_cast<int32_t>(f.obs_cam.size());
    pr.flags = refine_focal_length ? MSFM_BA_REFINE_FOCAL : 0;        // shared focal block of every residual (:225-233)
    pr.fx = f.fx; pr.fy = f.fy;
    pr.cams = f.cams.data(); pr.pts = f.pts.data(); pr.obs_uv = f.uv.data();
    pr.obs_cam = f.obs_cam.data(); pr.obs_pt = f.obs_pt.data(); pr.cam_const = f.cam_const.data();
    return pr;
}

msfm_ba* CreateOnDevice(const Flat& f, bool refine_focal_length) {
    const msfm_ba_problem pr = Describe(f, refine_focal_length);
    msfm_ba* ba = nullptr;
    device::Check(msfm_ba_create(device::Context(), &pr, &ba), "msfm_ba_create");
    return ba;
}
}  // namespace

double BundleData::Debug() {
    // mean over landmarks of the mean per-measurement reprojection error (BundleData.cpp:9-37); the per-measurement
    // error ||K [R|t] X - x|| equals the norm of the BA residual (Projection.cpp:114-133 vs CeresBundleOptimizer.cpp:44-51)
    Flat f = Flatten(*this);
    if (f.obs_cam.empty()) return 0.0;
    msfm_ba* ba = CreateOnDevice(f, false);
    std::vector<double> r(f.obs_cam.size() * 2);
    double cost = 0;
    device::Check(msfm_ba_evaluate(ba, r.data(), nullptr, &cost), "msfm_ba_evaluate");
    msfm_ba_destroy(ba);
    double sum = 0, num = 0;
    size_t i = 0;
    while (i < f.obs_pt.size()) {
        size_t e = i;
        double s = 0;
        while (e < f.obs_pt.size() && f.obs_pt[e] == f.obs_pt[i]) { s += std::sqrt(r[2 * e] * r[2 * e] + r[2 * e + 1] * r[2 * e + 1]); ++e; }
        sum += s / static_cast<double>(e - i);
        num += 1;
        i = e;
    }
    return sum / num;
}

CeresBundelOptimizer::CeresBundelOptimizer(const Parameters& params) : params_(params) {}

CeresBundelOptimizer::~CeresBundelOptimizer() {
    if (problem_ && device::Alive()) msfm_ba_destroy(problem_);      // the context's shutdown releases the device anyway
}

bool CeresBundelOptimizer::FilterStatistics(double max_reproj_error, std::vector<unsigned char>* obs_keep, std::vector<double>* pt_mean_error,
                                            std::vector<int>* pt_kept, std::vector<double>* pt_max_parallax_deg) {
    if (!problem_) return false;
    int64_t sizes[3];
    device::Check(msfm_ba_sizes(problem_, sizes), "msfm_ba_sizes");
    if (obs_keep) obs_keep->assign(static_cast<size_t>(sizes[2]), 0);
    if (pt_mean_error) pt_mean_error->assign(static_cast<size_t>(sizes[1]), 0.0);
    if (pt_kept) pt_kept->assign(static_cast<size_t>(sizes[1]), 0);
    if (pt_max_parallax_deg) pt_max_parallax_deg->assign(static_cast<size_t>(sizes[1]), 0.0);
    static_assert(sizeof(int) == sizeof(int32_t), "int32_t outputs are handed out as int");
    device::Check(msfm_ba_filter_stats(problem_, max_reproj_error, obs_keep ? obs_keep->data() : nullptr,
                                       pt_mean_error ? pt_mean_error->data() : nullptr,
                                       pt_kept ? reinterpret_cast<int32_t*>(pt_kept->data()) : nullptr,
                                       pt_max_parallax_deg ? pt_max_parallax_deg->data() : nullptr), "msfm_ba_filter_stats");
    return true;
}

bool CeresBundelOptimizer::Optimize(BundleData& bundle_data) {
    Flat f = Flatten(bundle_data);
    if (f.obs_cam.empty() || f.cam_ids.empty()) {
        std::cout << "Bundle Adjustment failed." << std::endl;
        return false;
    }
    // the device problem persists across calls: same sparsity pattern -> only the values travel; a changed map is analysed
    // again into the same device arena (msfm_ba_update)
    if (!problem_) {
        problem_ = CreateOnDevice(f, params_.refine_focal_length);
        last_reused_ = false;
    } else {
        const msfm_ba_problem pr = Describe(f, params_.refine_focal_length);
        int32_t reused = 0;
        device::Check(msfm_ba_update(problem_, &pr, &reused), "msfm_ba_update");
        last_reused_ = reused != 0;
    }
    msfm_ba* ba = problem_;
    {
        int64_t up[2] = {0, 0};
        device::Check(msfm_ba_last_upload(ba, up), "msfm_ba_last_upload");
        last_h2d_bytes_ = up[0];
    }
    msfm_ba_options opt;
    msfm_ba_default_options(&opt, static_cast<int32_t>(bundle_data.camera_poses.size()));   // :262-291
    msfm_ba_summary s;
    device::Check(msfm_ba_solve(ba, &opt, &s), "msfm_ba_solve");
    // parameters are written back in place whatever the outcome, like Ceres mutating the caller's blocks (:230-242)
    device::Check(msfm_ba_get_params(ba, f.cams.data(), f.pts.data()), "msfm_ba_get_params");
    double focal[2] = {f.fx, f.fy};
    device::Check(msfm_ba_get_focal(ba, focal), "msfm_ba_get_focal");
    for (size_t i = 0; i < f.cam_ids.size(); ++i) {
        BundleData::CameraPose& cp = bundle_data.camera_poses[f.cam_ids[i]];
        double* rv = PoseData(cp.rvec, "rvec");
        double* tv = PoseData(cp.tvec, "tvec");
        for (int k = 0; k < 3; ++k) {
            rv[k] = f.cams[6 * i + k];
            tv[k] = f.cams[6 * i + 3 + k];
        }
    }
    for (size_t p = 0; p < f.pt_ids.size(); ++p)
        for (int k = 0; k < 3; ++k) bundle_data.landmarks[f.pt_ids[p]].point3D(k) = f.pts[3 * p + k];
    last_iterations_ = s.iterations;
    last_initial_cost_ = s.initial_cost;
    last_final_cost_ = s.final_cost;
    if (s.termination != MSFM_BA_CONVERGENCE) {                                // :296-301
        std::cout << "Bundle Adjustment failed." << std::endl;
        return false;
    }
    std::cout << std::endl
              << "Bundle Adjustment statistics (approximated RMSE):\n"
              << " #residuals: " << s.num_residuals << "\n"
              << " Initial RMSE: " << std::sqrt(s.initial_cost * 2 / s.num_residuals) << "\n"
              << " Final RMSE: " << std::sqrt(s.final_cost * 2 / s.num_residuals) << "\n"
              << " Time (s): " << s.total_time_s << "\n"
              << std::endl;                                                    // :303-310
    if (params_.refine_focal_length) {                                         // :313-317
        bundle_data.K.at<double>(0, 0) = focal[0];
        bundle_data.K.at<double>(1, 1) = focal[1];
    }
    return true;
}
