// MonocularSfM::FeatureUtils — the three matching functions of the reference, same names / arguments / behaviour
// (reference: include/Feature/FeatureUtils.h:94-108, src/Feature/FeatureUtils.cpp:141-174,208-218,281-310),
// computed on the B200 through the C-ABI (include/msfm_b200.h).  Functions of the reference class that are not on
// the hot path (SIFT extraction, drawing) are intentionally absent.
#ifndef MSFM_HOST_FEATURE_UTILS_H_
#define MSFM_HOST_FEATURE_UTILS_H_
#include <vector>

#include "cvlite/cvlite.h"

namespace MonocularSfM {

class FeatureUtils {
public:
    // 1NN < distance_ratio * 2NN matches are kept; APPENDS to `matches` (FeatureUtils.cpp:154).
    static void ComputeMatches(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& matches,
                               const float distance_ratio = 0.8);
    // ComputeMatches both ways + CrossCheck (FeatureUtils.cpp:160-174), one distance matrix pass per direction on the GPU.
    static void ComputeCrossMatches(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& matches,
                                    const float distance_ratio = 0.8);
    // Host restatement incl. the unordered_map default-0 behaviour (FeatureUtils.cpp:281-310); the GPU path fuses it.
    static void CrossCheck(const std::vector<cv::DMatch>& matches12, const std::vector<cv::DMatch>& matches21,
                           std::vector<cv::DMatch>& prune_matches);
    static void FilterMatchesByDistance(const std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& prune_matches,
                                        const double& max_distance = 0.7);
    // Geometric verification (FeatureUtils.cpp:176-206, 220-233): keeps the matches that are inliers of a fundamental
    // matrix found by RANSAC (threshold 3 px, confidence 0.99).  The reference delegates to cv::findFundamentalMat; this is a
    // from-scratch estimator with the same error measure and stopping rule, not bit-compatible with OpenCV's sampler
    // (host/src/GeometricVerification.cpp).  Install with FeatureMatcher::SetGeometricFilter(FeatureUtils::FilterMatches).
    static void FilterMatches(const std::vector<cv::Point2f>& pts1, const std::vector<cv::Point2f>& pts2,
                              const std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& prune_matches);
    static void GetAlignedPointsFromMatches(const std::vector<cv::Point2f>& pts1, const std::vector<cv::Point2f>& pts2,
                                            const std::vector<cv::DMatch>& matches, std::vector<cv::Point2f>& aligned_pts1,
                                            std::vector<cv::Point2f>& aligned_pts2);
    // inlier mask of aligned correspondences; false (mask all zero) when no model with >= 8 inliers exists
    static bool FundamentalInliersRANSAC(const std::vector<cv::Point2f>& pts1, const std::vector<cv::Point2f>& pts2, double threshold,
                                         double confidence, std::vector<unsigned char>& mask);
    // Top-scale descriptor selection used by preemptive matching (FeatureUtils.cpp:68-96).
    static void ExtractTopScaleDescriptors(const std::vector<cv::KeyPoint> kpts, const cv::Mat& descriptors,
                                           const int& num_features, cv::Mat& top_scale_descriptors);

    // ---- bridge between the reference's CV_32F descriptors and the uint8 contract of the device path
    // CV_8U: used as is.  CV_32F with integral values in [0,255] (un-normalised SIFT): cast, lossless.
    // CV_32F otherwise (the reference's L1-root / L2 normalised rows): v -> clamp(round(512 v), 0, 255); this is the
    // COLMAP-style quantisation and is LOSSY w.r.t. the reference's float path (documented deviation).
    static cv::Mat ToUint8Descriptors(const cv::Mat& desc);
    // upload a CV_8U or CV_32F descriptor set under image_id; the float -> uint8 bridge runs on the device
    static void UploadDescriptors(int image_id, const cv::Mat& desc);
    // Distance scale of the quantised descriptors: a max_distance given for unit-norm floats is multiplied by this.
    static double QuantisationScale() { return 512.0; }
};

}  // namespace MonocularSfM
#endif
