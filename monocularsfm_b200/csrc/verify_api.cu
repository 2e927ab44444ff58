// C-ABI entry points of the batched geometric verification (include/msfm_b200.h): resident keypoints and
// msfm_verify_pairs / msfm_verify_pairs_dev.  Kernel: verify_ransac.cu.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace msfm {
cudaError_t launch_ransac_pairs(const void*, const int32_t*, const long long*, const int32_t*, int, double, double, int, uint8_t*,
                                int32_t*, int, cudaStream_t);
struct KpDevHost { const void* xy; int32_t n; int32_t pad; };
}  // namespace msfm
using namespace msfm;

extern "C" {

int msfm_keypoints_upload(msfm_ctx* c, int32_t image_id, const float* xy, int32_t n) {
    if (!c) return MSFM_E_INVALID;
    if (image_id < 0 || n < 0 || (n > 0 && !xy)) return c->fail(MSFM_E_INVALID, "msfm_keypoints_upload: bad arguments");
    MSFM_CUDA(c, cudaSetDevice(c->device));
    auto it = c->kps.find(image_id);
    if (it != c->kps.end()) {
        MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
        if (it->second.xy) cudaFree(it->second.xy);
        c->kps.erase(it);
    }
    msfm_ctx::KpHost k;
    k.n = n;
    if (n > 0) {
        MSFM_CUDA(c, cudaMalloc(&k.xy, static_cast<size_t>(n) * 2 * sizeof(float)));
        cudaError_t e = cudaMemcpyAsync(k.xy, xy, static_cast<size_t>(n) * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);          // the caller's buffer may be pageable and short-lived
        if (e != cudaSuccess) { cudaFree(k.xy); return c->cuda_fail(e, "msfm_keypoints_upload"); }
    }
    c->kps[image_id] = k;
    return MSFM_OK;
}

int msfm_keypoints_count(msfm_ctx* c, int32_t image_id) {
    if (!c) return MSFM_E_INVALID;
    auto it = c->kps.find(image_id);
    if (it == c->kps.end()) return c->fail(MSFM_E_NOT_FOUND, "keypoints of image %d not resident", image_id);
    return it->second.n;
}

int msfm_keypoints_release_all(msfm_ctx* c) {
    if (!c) return MSFM_E_INVALID;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& kv : c->kps)
        if (kv.second.xy) cudaFree(kv.second.xy);
    c->kps.clear();
    return MSFM_OK;
}

void msfm_verify_default_options(msfm_verify_options* o) {
    if (!o) return;
    o->threshold = 3.0;        // FeatureUtils.cpp:196: cv::findFundamentalMat(..., cv::FM_RANSAC, 3.0, 0.99, mask)
    o->confidence = 0.99;
    o->max_iters = 1000;       // OpenCV's default maxIters of that overload
    o->reserved = 0;
}

// pairs (host) -> per-pair table slots of the keypoint sets (device); validates the match indices' upper bounds lazily in
// the kernel's domain: indices come from msfm_match_pairs on descriptor sets of the same images.
static int build_tables(msfm_ctx* c, const int32_t* pairs, int32_t P, const char* who) {
    std::vector<KpDevHost> tab;
    std::vector<int32_t> slots(static_cast<size_t>(P) * 2);
    std::unordered_map<int32_t, int> slot_of;
    for (int p = 0; p < 2 * P; ++p) {
        const int32_t id = pairs[p];
        auto f = slot_of.find(id);
        if (f == slot_of.end()) {
            auto it = c->kps.find(id);
            if (it == c->kps.end()) return c->fail(MSFM_E_NOT_FOUND, "%s: keypoints of image %d not resident (msfm_keypoints_upload)", who, id);
            KpDevHost k; k.xy = it->second.xy; k.n = it->second.n; k.pad = 0;
            f = slot_of.emplace(id, static_cast<int>(tab.size())).first;
            tab.push_back(k);
        }
        slots[p] = f->second;
    }
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));        // tables of an earlier call may still be read
    MSFM_CUDA(c, c->d_kp_tab.reserve(std::max<size_t>(1, tab.size()) * sizeof(KpDevHost)));
    MSFM_CUDA(c, c->d_kp_slots.reserve(std::max<size_t>(1, slots.size()) * sizeof(int32_t)));
    if (!tab.empty()) MSFM_CUDA(c, cudaMemcpy(c->d_kp_tab.p, tab.data(), tab.size() * sizeof(KpDevHost), cudaMemcpyHostToDevice));
    if (!slots.empty()) MSFM_CUDA(c, cudaMemcpy(c->d_kp_slots.p, slots.data(), slots.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return MSFM_OK;
}

int msfm_verify_pairs_dev(msfm_ctx* c, const int32_t* pairs_host, int32_t P, const int64_t* offsets_dev, const int32_t* matches_dev,
                          const msfm_verify_options* opt, uint8_t* mask_dev, int32_t* counts_dev) {
    if (!c) return MSFM_E_INVALID;
    if (P < 0 || (P > 0 && (!pairs_host || !offsets_dev || !mask_dev)) || !opt || opt->reserved != 0 || !(opt->threshold > 0) ||
        !(opt->confidence > 0 && opt->confidence < 1) || opt->max_iters < 1)
        return c->fail(MSFM_E_INVALID, "msfm_verify_pairs: bad arguments");
    if (P == 0) return MSFM_OK;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    int rc = build_tables(c, pairs_host, P, "msfm_verify_pairs");
    if (rc) return rc;
    c->prof_begin(MSFM_PROF_VERIFY);
    MSFM_CUDA(c, launch_ransac_pairs(c->d_kp_tab.p, c->d_kp_slots.as<int32_t>(), reinterpret_cast<const long long*>(offsets_dev), matches_dev, P,
                                     opt->threshold, opt->confidence, opt->max_iters, mask_dev, counts_dev, c->num_sms, c->stream));
    c->prof_end();
    return MSFM_OK;
}

int msfm_verify_pairs(msfm_ctx* c, const int32_t* pairs, int32_t P, const int64_t* offsets, const int32_t* matches,
                      const msfm_verify_options* opt, uint8_t* inlier_mask, int32_t* inlier_counts) {
    if (!c) return MSFM_E_INVALID;
    if (P < 0 || (P > 0 && (!pairs || !offsets)) || !opt) return c->fail(MSFM_E_INVALID, "msfm_verify_pairs: bad arguments");
    if (P == 0) return MSFM_OK;
    const int64_t total = offsets[P];
    if (offsets[0] != 0 || total < 0 || (total > 0 && (!matches || !inlier_mask))) return c->fail(MSFM_E_INVALID, "msfm_verify_pairs: bad match lists");
    MSFM_CUDA(c, cudaSetDevice(c->device));
    // validate the match indices against the resident keypoint sets (host lists: cheap, and the kernel then needs no checks)
    for (int p = 0; p < P; ++p) {
        auto i1 = c->kps.find(pairs[2 * p]), i2 = c->kps.find(pairs[2 * p + 1]);
        if (i1 == c->kps.end() || i2 == c->kps.end())
            return c->fail(MSFM_E_NOT_FOUND, "msfm_verify_pairs: keypoints of image %d or %d not resident", pairs[2 * p], pairs[2 * p + 1]);
        if (offsets[p + 1] < offsets[p]) return c->fail(MSFM_E_INVALID, "msfm_verify_pairs: offsets must be non-decreasing");
        for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k)
            if (matches[2 * k] < 0 || matches[2 * k] >= i1->second.n || matches[2 * k + 1] < 0 || matches[2 * k + 1] >= i2->second.n)
                return c->fail(MSFM_E_INVALID, "msfm_verify_pairs: match %lld of pair %d indexes a keypoint that does not exist", (long long)k, p);
    }
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t off_bytes = (static_cast<size_t>(P) + 1) * sizeof(int64_t), m_bytes = static_cast<size_t>(total) * 2 * sizeof(int32_t);
    MSFM_CUDA(c, c->d_vf_in.reserve(off_bytes + m_bytes + 16));
    MSFM_CUDA(c, c->d_vf_out.reserve(static_cast<size_t>(total) + static_cast<size_t>(P) * sizeof(int32_t) + 16));
    int64_t* d_off = c->d_vf_in.as<int64_t>();
    int32_t* d_m = reinterpret_cast<int32_t*>(c->d_vf_in.as<uint8_t>() + off_bytes);
    int32_t* d_cnt = c->d_vf_out.as<int32_t>();
    uint8_t* d_mask = c->d_vf_out.as<uint8_t>() + static_cast<size_t>(P) * sizeof(int32_t);
    MSFM_CUDA(c, cudaMemcpyAsync(d_off, offsets, off_bytes, cudaMemcpyHostToDevice, c->stream));
    if (total > 0) MSFM_CUDA(c, cudaMemcpyAsync(d_m, matches, m_bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = msfm_verify_pairs_dev(c, pairs, P, d_off, d_m, opt, d_mask, d_cnt);
    if (rc) return rc;
    if (total > 0) MSFM_CUDA(c, cudaMemcpyAsync(inlier_mask, d_mask, static_cast<size_t>(total), cudaMemcpyDeviceToHost, c->stream));
    if (inlier_counts) MSFM_CUDA(c, cudaMemcpyAsync(inlier_counts, d_cnt, static_cast<size_t>(P) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    return MSFM_OK;
}

}  // extern "C"
