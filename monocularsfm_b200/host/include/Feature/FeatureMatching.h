// MonocularSfM::FeatureMatcher family — callers of the M-path (reference: include/Feature/FeatureMatching.h:28-141,
// src/Feature/FeatureMatching.cpp).  Same constructors, defaults and RunMatching loops; MatchImagePairs is re-plumbed:
// each image's descriptors are read from the database ONCE and kept resident on the GPU (the reference re-reads both
// blobs for every pair, FeatureMatching.cpp:31-33 "TODO cache"), and the whole batch of pairs is matched by ONE
// msfm_match_pairs call.
#ifndef MSFM_HOST_FEATURE_MATCHING_H_
#define MSFM_HOST_FEATURE_MATCHING_H_
#include <functional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "Common/Types.h"
#include "Database/Database.h"
#include "cvlite/cvlite.h"

namespace MonocularSfM {

class FeatureMatcher {
public:
    FeatureMatcher(const std::string& database_path, const int& max_num_matches = 10240, const double& max_distance = 0.7,
                   const double& distance_ratio = 0.8, const bool& cross_check = true)
        : database_path_(database_path), max_num_matches_(max_num_matches), max_distance_(max_distance),
          distance_ratio_(distance_ratio), cross_check_(cross_check), geometric_filter_(DefaultGeometricFilter()) {}
    virtual ~FeatureMatcher() {}

    void MatchImagePairs(const std::vector<std::pair<image_t, image_t>>& image_pairs);
    virtual void RunMatching() = 0;

    // Geometric verification (FeatureUtils::FilterMatches = F-matrix RANSAC, 3 px / 0.99, FeatureMatching.cpp:60).  As in the
    // reference it ALWAYS runs before WriteMatches.  By default the whole batch of pairs is verified on the device by ONE
    // msfm_verify_pairs call (keypoints are uploaded once per image).  SetGeometricFilter installs a per-pair HOST filter
    // instead — FeatureUtils::FilterMatches (DefaultGeometricFilter(), the host statement of the same estimator) or, in a
    // build with OpenCV, cv::findFundamentalMat; an empty function is an explicit opt-out (matches pass
    // through unverified).
    typedef std::function<void(const std::vector<cv::Point2f>&, const std::vector<cv::Point2f>&,
                               const std::vector<cv::DMatch>&, std::vector<cv::DMatch>&)> GeometricFilter;
    static GeometricFilter DefaultGeometricFilter();
    void SetGeometricFilter(GeometricFilter f) { geometric_filter_ = f; device_verification_ = false; }
    void SetVerbose(bool v) { verbose_ = v; }

protected:
    void EnsureResident(image_t image_id);
    std::string database_path_;
    int max_num_matches_;
    double max_distance_;
    double distance_ratio_;
    bool cross_check_;
    cv::Ptr<Database> database_;
    GeometricFilter geometric_filter_;
    bool device_verification_ = true;
    bool verbose_ = true;
    std::unordered_set<image_t> resident_;          // images whose descriptors (and keypoint positions) are on the device
    std::unordered_map<image_t, bool> quantised_;   // true when the CV_32F rows had to be quantised (x512)
};

class SequentialFeatureMatcher : public FeatureMatcher {
public:
    SequentialFeatureMatcher(const std::string& database_path, const int& overlap = 3, const int& max_num_matches = 10240,
                             const double& max_distance = 0.7, const double& distance_ratio = 0.8,
                             const bool& cross_check = true)
        : FeatureMatcher(database_path, max_num_matches, max_distance, distance_ratio, cross_check), overlap_(overlap) {}
    void RunMatching();
private:
    int overlap_;
};

class BruteFeatureMatcher : public FeatureMatcher {
public:
    BruteFeatureMatcher(const std::string& database_path, const int& max_pairs_size = 100, const bool& is_preemtive = true,
                        const int& preemtive_num_features = 100, const int& preemtive_min_num_matches = 4,
                        const int& max_num_matches = 10240, const double& max_distance = 0.7,
                        const double& distance_ratio = 0.8, const bool& cross_check = true)
        : FeatureMatcher(database_path, max_num_matches, max_distance, distance_ratio, cross_check),
          max_pairs_size_(max_pairs_size), is_preemtive_(is_preemtive), preemtive_num_features_(preemtive_num_features),
          preemtive_min_num_matches_(preemtive_min_num_matches) {}
    void RunMatching();
private:
    // Wu, "Towards Linear-Time Incremental Structure from Motion", 3DV 2013 (FeatureMatching.cpp:148-179): match the
    // preemtive_num_features_ largest-scale descriptors and keep pairs with >= preemtive_min_num_matches_ matches.
    std::vector<std::pair<image_t, image_t>> PreemptivelyFilterImagePairs(std::vector<std::pair<image_t, image_t>> image_pairs);
    void EnsureTopScaleResident(image_t image_id);
    int max_pairs_size_;
    bool is_preemtive_;
    int preemtive_num_features_;
    int preemtive_min_num_matches_;
    std::unordered_set<image_t> top_scale_resident_;
};

}  // namespace MonocularSfM
#endif
