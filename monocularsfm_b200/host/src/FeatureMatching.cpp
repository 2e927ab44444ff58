#include "Feature/FeatureMatching.h"

#include <algorithm>
#include <chrono>
#include <iostream>

#include "DeviceContext.h"
#include "Feature/FeatureUtils.h"

using namespace MonocularSfM;

namespace {
const int32_t kTopScaleIdBase = 1000000000;   // device ids of the preemptive-matching (top-scale) descriptor subsets
}

FeatureMatcher::GeometricFilter FeatureMatcher::DefaultGeometricFilter() { return FeatureUtils::FilterMatches; }

// Read an image's descriptors from the database once, bridge CV_32F -> u8, keep them resident on the device.
void FeatureMatcher::EnsureResident(image_t image_id) {
    if (resident_.count(image_id)) return;
    const cv::Mat desc = database_->ReadDescriptors(image_id);              // FeatureMatching.cpp:32-33
    msfm_ctx* ctx = device::Context();
    FeatureUtils::UploadDescriptors(image_id, desc);                        // CV_32F is bridged to uint8 on the device
    // integral float rows convert losslessly; anything else went through the x512 quantisation (synchronises: `desc`
    // may go out of scope)
    const int q = msfm_desc_quantised(ctx, image_id);
    if (q < 0) device::Check(q, "msfm_desc_quantised");
    quantised_[image_id] = q == 1;
    if (device_verification_) {
        // keypoint positions for the batched verification: read ONCE per image (the reference reads them per pair, :51-52)
        const std::vector<cv::KeyPoint> kpts = database_->ReadKeyPoints(image_id);
        std::vector<float> xy(kpts.size() * 2);
        for (size_t i = 0; i < kpts.size(); ++i) { xy[2 * i] = kpts[i].pt.x; xy[2 * i + 1] = kpts[i].pt.y; }
        device::Check(msfm_keypoints_upload(ctx, image_id, xy.data(), static_cast<int32_t>(kpts.size())), "msfm_keypoints_upload");
    }
    resident_.insert(image_id);
}

void FeatureMatcher::MatchImagePairs(const std::vector<std::pair<image_t, image_t>>& image_pairs) {
    database_->BeginTransaction();                                           // FeatureMatching.cpp:13
    const auto t0 = std::chrono::steady_clock::now();
    // pairs that already have a row are skipped — this is the reference's resume mechanism (:23-27)
    std::vector<std::pair<image_t, image_t>> todo;
    for (const auto& pr : image_pairs) {
        if (database_->ExistMatches(pr.first, pr.second)) {
            if (verbose_) std::cout << "Compute Matches " << pr.first << " - " << pr.second << " Existing, Continue!" << std::endl;
            continue;
        }
        todo.push_back(pr);
    }
    for (const auto& pr : todo) {
        EnsureResident(pr.first);
        EnsureResident(pr.second);
    }
    // FilterMatchesByDistance(max_distance_) (:49) compares distances on the scale of the descriptors: unit-norm floats that
    // went through the x512 quantisation need max_distance_ * 512, integral (un-normalised) sets the value as given.  A pair
    // mixing the two scales has no meaningful distance: quantise both sides the same way upstream.  The batch is split by
    // scale so that each device call carries ONE threshold.
    std::vector<std::pair<image_t, image_t>> all_todo;
    all_todo.swap(todo);
    for (int pass = 0; pass < 2; ++pass) {
    todo.clear();
    for (const auto& pr : all_todo) {
        const bool q1 = quantised_[pr.first], q2 = quantised_[pr.second];
        if (q1 != q2) {
            std::cerr << "FeatureMatcher: images " << pr.first << " and " << pr.second
                      << " hold descriptors on different scales (one set was quantised x512, the other is integral)" << std::endl;
            exit(EXIT_FAILURE);                                                  // the reference's error style (Database.cpp:16-20)
        }
        if ((q1 ? 1 : 0) == pass) todo.push_back(pr);
    }
    const double distance_scale = pass == 1 ? FeatureUtils::QuantisationScale() : 1.0;
    if (!todo.empty()) {
        std::vector<int32_t> ids(todo.size() * 2);
        int64_t bound = 0;
        msfm_ctx* ctx = device::Context();
        for (size_t p = 0; p < todo.size(); ++p) {
            ids[2 * p] = todo[p].first;
            ids[2 * p + 1] = todo[p].second;
            bound += msfm_desc_count(ctx, todo[p].first);
        }
        msfm_match_options opt;
        opt.max_distance = max_distance_ < 0 ? -1.0 : max_distance_ * distance_scale;
        opt.distance_ratio = static_cast<float>(distance_ratio_);            // narrowed like FeatureUtils.h:95
        opt.cross_check = cross_check_ ? 1 : 0;
        opt.opencv_quirks = 1;
        opt.reserved = 0;
        std::vector<int64_t> offsets(todo.size() + 1);
        std::vector<int32_t> out(static_cast<size_t>(std::max<int64_t>(1, bound)) * 2);
        std::vector<float> dist(std::max<int64_t>(1, bound));
        int64_t total = 0;
        device::Check(msfm_match_pairs(ctx, ids.data(), static_cast<int32_t>(todo.size()), &opt, offsets.data(), out.data(),
                                       dist.data(), bound, &total), "msfm_match_pairs");
        // FeatureUtils::FilterMatches (:60) for the whole batch in one device call
        std::vector<uint8_t> inlier(static_cast<size_t>(std::max<int64_t>(1, total)), 1);
        if (device_verification_) {
            msfm_verify_options vo;
            msfm_verify_default_options(&vo);                                   // 3.0 px, 0.99 (FeatureUtils.cpp:196)
            device::Check(msfm_verify_pairs(ctx, ids.data(), static_cast<int32_t>(todo.size()), offsets.data(), out.data(), &vo,
                                            inlier.data(), nullptr), "msfm_verify_pairs");
        }
        for (size_t p = 0; p < todo.size(); ++p) {
            std::vector<cv::DMatch> prune_matches;
            for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k)
                if (inlier[k]) prune_matches.push_back(cv::DMatch(out[2 * k], out[2 * k + 1], 0, static_cast<float>(dist[k] / distance_scale)));
            std::vector<cv::DMatch> verified;
            if (!device_verification_ && geometric_filter_) {
                std::vector<cv::KeyPoint> k1 = database_->ReadKeyPoints(todo[p].first), k2 = database_->ReadKeyPoints(todo[p].second);
                std::vector<cv::Point2f> p1(k1.size()), p2(k2.size());
                for (size_t i = 0; i < k1.size(); ++i) p1[i] = k1[i].pt;
                for (size_t i = 0; i < k2.size(); ++i) p2[i] = k2[i].pt;
                geometric_filter_(p1, p2, prune_matches, verified);          // FeatureUtils::FilterMatches (:60)
            } else {
                verified.swap(prune_matches);
            }
            if (verbose_) {
                std::cout << "Compute Matches " << todo[p].first << " - " << todo[p].second << " ... " << std::endl;
                std::cout << "\t matches num : " << verified.size() << std::endl;
            }
            database_->WriteMatches(todo[p].first, todo[p].second, verified);   // a row is written even for 0 matches (:68-70)
        }
    }
    }   // scale passes
    todo.swap(all_todo);
    if (verbose_ && !todo.empty()) {
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "\t batch of " << todo.size() << " pairs: " << s << " s" << std::endl;
    }
    database_->EndTransaction();                                             // :72
}

void SequentialFeatureMatcher::RunMatching() {
    database_ = cv::Ptr<Database>(new Database());
    database_->Open(database_path_);
    std::vector<Database::Image> images = database_->ReadAllImages();
    // each image against its `overlap_` predecessors (FeatureMatching.cpp:82-98); the image list index is used as id,
    // exactly as the reference does
    for (size_t i = 1; i < images.size(); ++i) {
        std::vector<std::pair<image_t, image_t>> image_pairs;
        for (int k = 1; k <= overlap_; ++k) {
            const int j = static_cast<int>(i) - k;
            if (j < 0) break;
            image_pairs.push_back(std::make_pair(static_cast<image_t>(i), static_cast<image_t>(j)));
        }
        MatchImagePairs(image_pairs);
    }
    database_->Close();
}

void BruteFeatureMatcher::RunMatching() {
    database_ = cv::Ptr<Database>(new Database());
    database_->Open(database_path_);
    std::vector<Database::Image> images = database_->ReadAllImages();
    // The pair order is the reference's (FeatureMatching.cpp:110-142: for i, for j < i).  The reference also flushes at the end
    // of every i because its batch is a list of descriptor Mats loaded for that image; here the descriptors are resident on the
    // device and a batch is ONE device call (matching + verification) and one transaction, so batches are filled up to
    // max_pairs_size across images: 32 images give 5 batches of <= 100 pairs instead of 31 batches of 1..31.
    std::vector<std::pair<image_t, image_t>> image_pairs;
    auto flush = [&]() {
        if (image_pairs.empty()) return;
        if (is_preemtive_) image_pairs = PreemptivelyFilterImagePairs(image_pairs);
        MatchImagePairs(image_pairs);
        image_pairs.clear();
    };
    for (size_t i = 0; i < images.size(); ++i)
        for (size_t j = 0; j < i; ++j) {
            image_pairs.push_back(std::make_pair(static_cast<image_t>(i), static_cast<image_t>(j)));
            if (static_cast<int>(image_pairs.size()) >= std::max(1, max_pairs_size_)) flush();
        }
    flush();
    database_->Close();
}

void BruteFeatureMatcher::EnsureTopScaleResident(image_t image_id) {
    if (top_scale_resident_.count(image_id)) return;
    const std::vector<cv::KeyPoint> kpts = database_->ReadKeyPoints(image_id);
    const cv::Mat desc = database_->ReadDescriptors(image_id);
    cv::Mat top;
    FeatureUtils::ExtractTopScaleDescriptors(kpts, desc, preemtive_num_features_, top);   // :180-196
    msfm_ctx* ctx = device::Context();
    FeatureUtils::UploadDescriptors(kTopScaleIdBase + image_id, top);
    device::Check(msfm_sync(ctx), "msfm_sync");
    top_scale_resident_.insert(image_id);
}

std::vector<std::pair<image_t, image_t>> BruteFeatureMatcher::PreemptivelyFilterImagePairs(
    std::vector<std::pair<image_t, image_t>> image_pairs) {
    std::vector<std::pair<image_t, image_t>> filtered;
    if (image_pairs.empty()) return filtered;
    std::vector<int32_t> ids(image_pairs.size() * 2);
    for (size_t p = 0; p < image_pairs.size(); ++p) {
        EnsureTopScaleResident(image_pairs[p].first);
        EnsureTopScaleResident(image_pairs[p].second);
        ids[2 * p] = kTopScaleIdBase + image_pairs[p].first;
        ids[2 * p + 1] = kTopScaleIdBase + image_pairs[p].second;
    }
    msfm_match_options opt;
    opt.max_distance = -1.0;                                                 // the pre-pass applies no distance filter (:160-170)
    opt.distance_ratio = static_cast<float>(distance_ratio_);
    opt.cross_check = cross_check_ ? 1 : 0;
    opt.opencv_quirks = 1;
    opt.reserved = 0;
    const int64_t cap = static_cast<int64_t>(image_pairs.size()) * std::max(1, preemtive_num_features_);
    std::vector<int64_t> offsets(image_pairs.size() + 1);
    std::vector<int32_t> out(static_cast<size_t>(cap) * 2);
    int64_t total = 0;
    device::Check(msfm_match_pairs(device::Context(), ids.data(), static_cast<int32_t>(image_pairs.size()), &opt, offsets.data(),
                                   out.data(), nullptr, cap, &total), "msfm_match_pairs(preemptive)");
    for (size_t p = 0; p < image_pairs.size(); ++p)
        if (offsets[p + 1] - offsets[p] >= preemtive_min_num_matches_) filtered.push_back(image_pairs[p]);   // :172-173
    return filtered;
}
