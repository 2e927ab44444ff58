// MonocularSfM::BundleData — BA problem container, unchanged API (reference: include/Optimizer/BundleData.h:19-65).
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
    struct Measurement {
        image_t image_id;
        cv::Vec2d point2D;
        Measurement(const image_t& image_id, const cv::Vec2d& point2D) : image_id(image_id), point2D(point2D) {}
    };
    struct Landmark {
        cv::Vec3d point3D;
        std::vector<Measurement> measurements;
        Landmark() {}
        Landmark(const cv::Vec3d& point3D, const std::vector<Measurement>& measurements)
            : point3D(point3D), measurements(measurements) {}
    };
    struct CameraPose {
        cv::Mat rvec;   // 3x1 CV_64F
        cv::Mat tvec;   // 3x1 CV_64F
        CameraPose() {}
        CameraPose(const cv::Mat& rvec, const cv::Mat& tvec) : rvec(rvec), tvec(tvec) {}
    };
    cv::Mat K;          // 3x3 CV_64F
    std::unordered_map<point3D_t, Landmark> landmarks;
    std::unordered_map<image_t, CameraPose> camera_poses;
    std::unordered_set<image_t> constant_camera_pose;

    // mean over landmarks of the mean reprojection error (pixels) — BundleData.cpp:9-37, evaluated on the device.
    double Debug();
};

}  // namespace MonocularSfM
#endif
