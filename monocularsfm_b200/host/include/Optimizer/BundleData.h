// MonocularSfM::BundleData — the BA problem container callers fill (Map::GetLocalBAData / GetGlobalBAData,
// src/Reconstruction/Map.cpp:965-1173) and CeresBundelOptimizer::Optimize mutates in place.  The member names and types
// are the reference's (include/Optimizer/BundleData.h:19-65) because callers touch them directly; everything below the
// type level is different: Optimize flattens the hash maps into the SoA arrays of include/msfm_b200.h
// (host/src/BundleOptimizer.cpp: Flatten):
//
//   camera_poses    -> cams [n_cams][6] = rvec | tvec, in ascending image_id order (hash order is unspecified in the
//                      reference; a fixed order makes runs reproducible)
//   landmarks       -> pts [n_pts][3] in ascending point3D_t order, and one observation per Measurement, grouped by
//                      landmark: obs_pt (non-decreasing), obs_cam (index into cams), obs_uv = point2D - (cx, cy)
//                      (the centring of CeresBundleOptimizer.cpp:221-222)
//   constant_camera_pose -> cam_const [n_cams] (SetParameterBlockConstant, :256-260)
//   K               -> fx = K(0,0), fy = K(1,1), cx = K(0,2), cy = K(1,2) (:197-200); K(0,0), K(1,1) are written back
//                      when Parameters::refine_focal_length is set (:313-317)
#ifndef MSFM_HOST_BUNDLE_DATA_H_
#define MSFM_HOST_BUNDLE_DATA_H_
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "Common/Types.h"
#include "cvlite/cvlite.h"

namespace MonocularSfM {

class BundleData {
public:
    // one 2-D observation of a landmark: pixel coordinates in image `image_id` (undistorted once at load, Map.cpp:98-103)
    struct Measurement {
        Measurement(const image_t& image_id, const cv::Vec2d& point2D) : image_id(image_id), point2D(point2D) {}
        image_t image_id;
        cv::Vec2d point2D;
    };
    // a 3-D point and its track; a landmark whose measurements all belong to constant cameras still contributes residuals
    struct Landmark {
        Landmark() {}
        Landmark(const cv::Vec3d& point3D, const std::vector<Measurement>& measurements)
            : point3D(point3D), measurements(measurements) {}
        cv::Vec3d point3D;
        std::vector<Measurement> measurements;
    };
    // world -> camera: x_cam = R(rvec) x + tvec, rvec an angle-axis vector (cv::Rodrigues); both CV_64F with 3 elements (3x1 or 1x3)
    struct CameraPose {
        CameraPose() {}
        CameraPose(const cv::Mat& rvec, const cv::Mat& tvec) : rvec(rvec), tvec(tvec) {}
        cv::Mat rvec;
        cv::Mat tvec;
    };

    cv::Mat K;                                                  // 3x3 CV_64F, shared by all cameras
    std::unordered_map<point3D_t, Landmark> landmarks;
    std::unordered_map<image_t, CameraPose> camera_poses;
    std::unordered_set<image_t> constant_camera_pose;           // poses the optimizer must not move

    // mean over landmarks of the mean reprojection error in pixels (BundleData.cpp:9-37).  Evaluated on the device with the
    // residual kernel of the optimizer: ||K [R|t] X - x|| equals the norm of the BA residual (Projection.cpp:114-133).
    double Debug();
};

}  // namespace MonocularSfM
#endif
