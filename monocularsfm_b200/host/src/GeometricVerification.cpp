// FeatureUtils::FilterMatches / GetAlignedPointsFromMatches — geometric verification of a pair's matches.
//
// The reference calls cv::findFundamentalMat(aligned_pts1, aligned_pts2, cv::FM_RANSAC, 3.0, 0.99, inlier_mask)
// (src/Feature/FeatureUtils.cpp:176-206) and keeps the matches whose mask byte is set.  This file is a from-scratch
// statement of that estimator for builds without OpenCV (a build with -DMSFM_WITH_OPENCV can install
// cv::findFundamentalMat itself through FeatureMatcher::SetGeometricFilter): RANSAC over minimal 8-point samples with
// Hartley normalisation, the rank-2 constraint, OpenCV's error measure (the larger of the two squared point-to-epipolar-
// line distances against threshold^2) and its adaptive iteration count log(1 - confidence) / log(1 - w^8), at most 1000,
// followed by refits on the consensus set while they explain more points.
// It is NOT bit-compatible with OpenCV's RANSAC (different sampler, 8- instead of 7-point minimal solver): the inlier
// sets agree where the geometry is unambiguous, which is what tests/test_host_cpp.py checks against cv2 itself.
// CPU code: the matches of one pair are at most a few thousand points (SURVEY.md §8f-1 ranks a batched GPU version next).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "Feature/FeatureUtils.h"

using namespace MonocularSfM;

namespace {

// eigen-decomposition of a symmetric n x n matrix (row-major, n <= 9) by cyclic Jacobi rotations;
// eigenvalues in w, eigenvectors in the COLUMNS of v
void JacobiEigen(int n, double* a, double* w, double* v) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) v[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
        if (off < 1e-30) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = c * akp - s * akq;
                    a[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = c * apk - s * aqk;
                    a[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = v[k * n + p], vkq = v[k * n + q];
                    v[k * n + p] = c * vkp - s * vkq;
                    v[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
}

struct Normaliser {
    double cx = 0, cy = 0, s = 1;
    void fit(const std::vector<cv::Point2f>& p, const std::vector<int>& idx) {
        cx = cy = 0;
        for (int i : idx) { cx += p[i].x; cy += p[i].y; }
        cx /= idx.size(); cy /= idx.size();
        double d = 0;
        for (int i : idx) d += std::sqrt((p[i].x - cx) * (p[i].x - cx) + (p[i].y - cy) * (p[i].y - cy));
        d /= idx.size();
        s = d > 1e-12 ? std::sqrt(2.0) / d : 1.0;
    }
};

// Eight-point algorithm on the correspondences `idx` (>= 8): x2^T F x1 = 0.  Returns false for a degenerate sample.
bool EightPoint(const std::vector<cv::Point2f>& p1, const std::vector<cv::Point2f>& p2, const std::vector<int>& idx, double F[9]) {
    Normaliser n1, n2;
    n1.fit(p1, idx);
    n2.fit(p2, idx);
    double ata[81];
    for (double& x : ata) x = 0.0;
    for (int i : idx) {
        const double x1 = (p1[i].x - n1.cx) * n1.s, y1 = (p1[i].y - n1.cy) * n1.s;
        const double x2 = (p2[i].x - n2.cx) * n2.s, y2 = (p2[i].y - n2.cy) * n2.s;
        const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
        for (int a = 0; a < 9; ++a)
            for (int b = 0; b < 9; ++b) ata[a * 9 + b] += r[a] * r[b];
    }
    double w[9], v[81];
    JacobiEigen(9, ata, w, v);
    int k = 0;
    for (int i = 1; i < 9; ++i)
        if (w[i] < w[k]) k = i;
    double f[9];
    for (int i = 0; i < 9; ++i) f[i] = v[i * 9 + k];
    // rank 2: F = U diag(s1, s2, 0) V^T through the eigen-decomposition of F^T F
    double ftf[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) ftf[a * 3 + b] = f[0 * 3 + a] * f[0 * 3 + b] + f[1 * 3 + a] * f[1 * 3 + b] + f[2 * 3 + a] * f[2 * 3 + b];
    double e[3], V[9];
    JacobiEigen(3, ftf, e, V);
    int lo = 0;
    for (int i = 1; i < 3; ++i)
        if (e[i] < e[lo]) lo = i;
    // remove the component along the weakest right singular vector:  F <- F (I - v v^T)
    const double vx = V[0 * 3 + lo], vy = V[1 * 3 + lo], vz = V[2 * 3 + lo];
    double fr[9];
    for (int a = 0; a < 3; ++a) {
        const double d = f[a * 3] * vx + f[a * 3 + 1] * vy + f[a * 3 + 2] * vz;
        fr[a * 3] = f[a * 3] - d * vx; fr[a * 3 + 1] = f[a * 3 + 1] - d * vy; fr[a * 3 + 2] = f[a * 3 + 2] - d * vz;
    }
    // denormalise: F = T2^T Fn T1,  T = [s 0 -s cx; 0 s -s cy; 0 0 1]
    const double T1[9] = {n1.s, 0, -n1.s * n1.cx, 0, n1.s, -n1.s * n1.cy, 0, 0, 1};
    const double T2[9] = {n2.s, 0, -n2.s * n2.cx, 0, n2.s, -n2.s * n2.cy, 0, 0, 1};
    double tmp[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) tmp[a * 3 + b] = fr[a * 3] * T1[b] + fr[a * 3 + 1] * T1[3 + b] + fr[a * 3 + 2] * T1[6 + b];
    double nrm = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            F[a * 3 + b] = T2[a] * tmp[b] + T2[3 + a] * tmp[3 + b] + T2[6 + a] * tmp[6 + b];
            nrm += F[a * 3 + b] * F[a * 3 + b];
        }
    if (!(nrm > 1e-300) || !std::isfinite(nrm)) return false;
    nrm = 1.0 / std::sqrt(nrm);
    for (int i = 0; i < 9; ++i) F[i] *= nrm;
    return true;
}

// OpenCV's error of a correspondence under F: the larger of the squared distances of x2 to the line F x1 and of x1 to
// the line F^T x2
inline double EpipolarError(const double F[9], const cv::Point2f& a, const cv::Point2f& b) {
    const double l0 = F[0] * a.x + F[1] * a.y + F[2], l1 = F[3] * a.x + F[4] * a.y + F[5], l2 = F[6] * a.x + F[7] * a.y + F[8];
    const double d2 = b.x * l0 + b.y * l1 + l2;
    const double e2 = d2 * d2 / (l0 * l0 + l1 * l1);
    const double m0 = F[0] * b.x + F[3] * b.y + F[6], m1 = F[1] * b.x + F[4] * b.y + F[7], m2 = F[2] * b.x + F[5] * b.y + F[8];
    const double d1 = a.x * m0 + a.y * m1 + m2;
    const double e1 = d1 * d1 / (m0 * m0 + m1 * m1);
    return std::max(e1, e2);
}

int CountInliers(const double F[9], const std::vector<cv::Point2f>& p1, const std::vector<cv::Point2f>& p2, double thr2,
                 std::vector<unsigned char>* mask) {
    int n = 0;
    for (size_t i = 0; i < p1.size(); ++i) {
        const double e = EpipolarError(F, p1[i], p2[i]);
        const bool in = e <= thr2;            // NaN (degenerate line) compares false
        if (mask) (*mask)[i] = in ? 1 : 0;
        n += in ? 1 : 0;
    }
    return n;
}

}  // namespace

void FeatureUtils::GetAlignedPointsFromMatches(const std::vector<cv::Point2f>& pts1, const std::vector<cv::Point2f>& pts2,
                                               const std::vector<cv::DMatch>& matches, std::vector<cv::Point2f>& aligned_pts1,
                                               std::vector<cv::Point2f>& aligned_pts2) {
    for (const cv::DMatch& m : matches) {          // FeatureUtils.cpp:220-233: queryIdx indexes pts1, trainIdx pts2
        aligned_pts1.push_back(pts1[m.queryIdx]);
        aligned_pts2.push_back(pts2[m.trainIdx]);
    }
}

bool FeatureUtils::FundamentalInliersRANSAC(const std::vector<cv::Point2f>& p1, const std::vector<cv::Point2f>& p2, double threshold,
                                            double confidence, std::vector<unsigned char>& mask) {
    const int n = static_cast<int>(p1.size());
    mask.assign(n, 0);
    if (n < 8 || p2.size() != p1.size()) return false;
    const double thr2 = threshold * threshold;
    uint64_t rng = 0x9E3779B97F4A7C15ull;                     // fixed seed: the result is a function of the input only
    auto next = [&rng]() {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;  // xorshift64
        return rng;
    };
    int best = 0, max_iters = 1000;
    double Fbest[9] = {0};
    std::vector<int> sample(8);
    double F[9];
    for (int it = 0; it < max_iters; ++it) {
        for (int k = 0; k < 8;) {                              // 8 distinct indices
            const int c = static_cast<int>(next() % static_cast<uint64_t>(n));
            bool dup = false;
            for (int j = 0; j < k; ++j) dup = dup || sample[j] == c;
            if (!dup) sample[k++] = c;
        }
        if (!EightPoint(p1, p2, sample, F)) continue;
        const int cnt = CountInliers(F, p1, p2, thr2, nullptr);
        if (cnt > best) {
            best = cnt;
            for (int i = 0; i < 9; ++i) Fbest[i] = F[i];
            // adaptive stop: enough samples to have drawn an all-inlier one with the requested confidence
            const double w8 = std::pow(static_cast<double>(cnt) / n, 8);
            if (w8 >= 1.0) {
                max_iters = it + 1;                                   // every point is an inlier
            } else if (w8 > 1e-12) {                                  // below that the bound exceeds any iteration budget
                const double need = std::log(1.0 - confidence) / std::log1p(-w8);
                if (need < static_cast<double>(max_iters)) max_iters = std::max(it + 1, static_cast<int>(std::ceil(need)));
            }
        }
    }
    if (best < 8) return false;
    CountInliers(Fbest, p1, p2, thr2, &mask);
    // local optimisation: a minimal sample of noisy points gives a model that misses true correspondences near the
    // threshold; refit on the consensus set and keep the refit while it explains more points
    for (int round = 0; round < 4; ++round) {
        std::vector<int> in;
        for (int i = 0; i < n; ++i)
            if (mask[i]) in.push_back(i);
        if (in.size() < 8 || !EightPoint(p1, p2, in, F)) break;
        std::vector<unsigned char> m2(n, 0);
        const int cnt = CountInliers(F, p1, p2, thr2, &m2);
        if (cnt <= best) break;
        best = cnt;
        mask.swap(m2);
    }
    return true;
}

void FeatureUtils::FilterMatches(const std::vector<cv::Point2f>& pts1, const std::vector<cv::Point2f>& pts2,
                                 const std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& prune_matches) {
    if (pts1.empty() || matches.empty()) return;                                   // FeatureUtils.cpp:181-184
    std::vector<cv::Point2f> a1, a2;
    GetAlignedPointsFromMatches(pts1, pts2, matches, a1, a2);
    if (a1.empty()) return;
    std::vector<unsigned char> mask;
    if (!FundamentalInliersRANSAC(a1, a2, 3.0, 0.99, mask)) return;               // :196 (FM_RANSAC, 3.0, 0.99)
    for (size_t i = 0; i < mask.size(); ++i)
        if (mask[i]) prune_matches.push_back(matches[i]);                          // :198-204
}
