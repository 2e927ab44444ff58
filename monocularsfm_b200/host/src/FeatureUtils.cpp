#include "Feature/FeatureUtils.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <unordered_map>

#include "DeviceContext.h"

using namespace MonocularSfM;

namespace {
const int32_t kScratchId1 = 2000000001, kScratchId2 = 2000000002;   // image ids reserved for the two-Mat entry points

void MatchTwoMats(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& matches, float distance_ratio,
                  bool cross_check) {
    msfm_ctx* ctx = device::Context();
    FeatureUtils::UploadDescriptors(kScratchId1, desc1);
    device::Check(msfm_sync(ctx), "msfm_sync");   // the upload staging buffer is reused by the next upload
    FeatureUtils::UploadDescriptors(kScratchId2, desc2);
    // Both operands must live on ONE scale.  If either float set had to be quantised (x512), the other is re-uploaded with
    // forced quantisation, and the distances handed back are divided by the scale: a caller that keeps the reference idiom
    // ComputeCrossMatches(...) + FilterMatchesByDistance(m, out, 0.7) sees unit-scale L2 distances as with OpenCV.
    // Saturation: components above 255/512 = 0.498 clamp to 255; for such (sparse RootSIFT-like) rows distances and the
    // ratio test can differ from the float path — documented deviation, covered by the host tests.
    int q1 = msfm_desc_quantised(ctx, kScratchId1), q2 = msfm_desc_quantised(ctx, kScratchId2);
    if (q1 < 0) device::Check(q1, "msfm_desc_quantised");
    if (q2 < 0) device::Check(q2, "msfm_desc_quantised");
    if (q1 != q2) {
        const cv::Mat& other = q1 ? desc2 : desc1;
        if (other.type() == CV_32F) {
            device::Check(msfm_desc_upload_f32(ctx, q1 ? kScratchId2 : kScratchId1, reinterpret_cast<const float*>(other.data), other.rows, 1),
                          "msfm_desc_upload_f32");
            device::Check(msfm_sync(ctx), "msfm_sync");
        }
    }
    const double scale = (q1 || q2) ? FeatureUtils::QuantisationScale() : 1.0;
    const cv::Mat& a = desc1;
    msfm_match_options opt;
    opt.max_distance = -1.0;
    opt.distance_ratio = distance_ratio;
    opt.cross_check = cross_check ? 1 : 0;
    opt.opencv_quirks = 1;
    opt.reserved = 0;
    const int32_t pair[2] = {kScratchId1, kScratchId2};
    const int64_t cap = std::max(1, a.rows);
    std::vector<int32_t> out(static_cast<size_t>(cap) * 2);
    std::vector<float> dist(cap);
    int64_t offsets[2] = {0, 0}, total = 0;
    device::Check(msfm_match_pairs(ctx, pair, 1, &opt, offsets, out.data(), dist.data(), cap, &total), "msfm_match_pairs");
    for (int64_t k = 0; k < total; ++k)
        matches.push_back(cv::DMatch(out[2 * k], out[2 * k + 1], 0, static_cast<float>(dist[k] / scale)));   // imgIdx 0 like knnMatch on one train set
}
}  // namespace

// CV_8U rows go to the device as they are; CV_32F rows (what Database::ReadDescriptors returns, Database.cpp:510-523) are
// copied as float and bridged to uint8 ON THE DEVICE by the rule ToUint8Descriptors states on the host.
void FeatureUtils::UploadDescriptors(int image_id, const cv::Mat& desc) {
    assert(desc.cols == MSFM_DESC_DIM || desc.rows == 0);
    assert(desc.type() == CV_8U || desc.type() == CV_32F);
    msfm_ctx* ctx = device::Context();
    if (desc.type() == CV_8U) device::Check(msfm_desc_upload_u8(ctx, image_id, desc.data, desc.rows), "msfm_desc_upload_u8");
    else device::Check(msfm_desc_upload_f32(ctx, image_id, reinterpret_cast<const float*>(desc.data), desc.rows, 0), "msfm_desc_upload_f32");
}

cv::Mat FeatureUtils::ToUint8Descriptors(const cv::Mat& desc) {
    if (desc.type() == CV_8U) {
        assert(desc.cols == MSFM_DESC_DIM || desc.rows == 0);
        return desc;
    }
    assert(desc.type() == CV_32F);
    assert(desc.cols == MSFM_DESC_DIM || desc.rows == 0);
    cv::Mat out(desc.rows, MSFM_DESC_DIM, CV_8U);
    bool integral = true;
    for (int i = 0; i < desc.rows && integral; ++i) {
        const float* r = desc.ptr<float>(i);
        for (int j = 0; j < desc.cols; ++j) {
            const float v = r[j];
            if (!(v >= 0.f && v <= 255.f && v == std::floor(v))) { integral = false; break; }
        }
    }
    for (int i = 0; i < desc.rows; ++i) {
        const float* r = desc.ptr<float>(i);
        unsigned char* o = out.ptr<unsigned char>(i);
        for (int j = 0; j < desc.cols; ++j) {
            const float v = integral ? r[j] : std::nearbyint(r[j] * 512.0f);
            o[j] = static_cast<unsigned char>(std::min(255.f, std::max(0.f, v)));
        }
    }
    return out;
}

void FeatureUtils::ComputeMatches(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& matches,
                                  const float distance_ratio) {
    MatchTwoMats(desc1, desc2, matches, distance_ratio, false);
}

void FeatureUtils::ComputeCrossMatches(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& matches,
                                       const float distance_ratio) {
    MatchTwoMats(desc1, desc2, matches, distance_ratio, true);
}

void FeatureUtils::CrossCheck(const std::vector<cv::DMatch>& matches12, const std::vector<cv::DMatch>& matches21,
                              std::vector<cv::DMatch>& prune_matches) {
    // reverse lookup; operator[] on a missing key yields 0, which the reference compares against queryIdx (:302)
    std::unordered_map<int, int> reverse;
    for (const cv::DMatch& m : matches21) reverse[m.queryIdx] = m.trainIdx;
    for (const cv::DMatch& m : matches12)
        if (reverse[m.trainIdx] == m.queryIdx) prune_matches.push_back(m);
}

void FeatureUtils::FilterMatchesByDistance(const std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& prune_matches,
                                           const double& max_distance) {
    for (const cv::DMatch& m : matches)
        if (!(m.distance > max_distance)) prune_matches.push_back(m);
}

void FeatureUtils::ExtractTopScaleDescriptors(const std::vector<cv::KeyPoint> kpts, const cv::Mat& descriptors,
                                              const int& num_features, cv::Mat& top_scale_descriptors) {
    if (num_features > static_cast<int>(kpts.size())) {
        top_scale_descriptors = descriptors;
        return;
    }
    std::vector<std::pair<size_t, float>> scales;
    scales.reserve(kpts.size());
    for (size_t i = 0; i < kpts.size(); ++i) scales.emplace_back(i, kpts[i].size);
    std::partial_sort(scales.begin(), scales.begin() + num_features, scales.end(),
                      [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) { return a.second > b.second; });
    top_scale_descriptors = cv::Mat(num_features, descriptors.cols, descriptors.type());
    const size_t row_bytes = static_cast<size_t>(descriptors.cols) * descriptors.elemSize();
    for (int i = 0; i < num_features; ++i)
        std::memcpy(top_scale_descriptors.data + i * row_bytes, descriptors.data + scales[i].first * row_bytes, row_bytes);
}
