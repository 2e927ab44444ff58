"""M-path oracle (TEST INFRASTRUCTURE — never imported by the product path).

CPU restatement of the reference's descriptor matching path:

* ``knn2``                     restates OpenCV ``BFMatcher(NORM_L2).knnMatch(q, t, k=2)`` as the reference
                               calls it in ``src/Feature/FeatureUtils.cpp:146-149``.  The arithmetic lives in
                               third-party OpenCV (not vendored under /root/reference; README pins "3.x or
                               higher"; this container has opencv-python-headless 4.13.0).  Published algorithm
                               (modules/core/src/batch_distance.cpp): ``dist[j] = sqrtf(sum_k (q_k - t_jk)^2)``
                               accumulated exactly (integers < 2^24), then a stable insertion top-K on the
                               float distance: an element is inserted only if ``d < dist[K-1]`` and lands after
                               every element with ``dist <= d``  ==> lexicographic (float distance, index) order.
* ``compute_matches``          FeatureUtils::ComputeMatches       src/Feature/FeatureUtils.cpp:141-157
* ``cross_check``              FeatureUtils::CrossCheck           src/Feature/FeatureUtils.cpp:281-310
                               (including the ``unordered_map::operator[]`` default-0 quirk)
* ``compute_cross_matches``    FeatureUtils::ComputeCrossMatches  src/Feature/FeatureUtils.cpp:160-174
* ``filter_matches_by_distance`` FeatureUtils::FilterMatchesByDistance src/Feature/FeatureUtils.cpp:208-218
* ``match_image_pair``         the descriptor part of FeatureMatcher::MatchImagePairs
                               src/Feature/FeatureMatching.cpp:36-49

Pinning: ``tests/test_oracle_match.py`` checks ``knn2`` against cv2 itself (when importable) and against the
committed golden vectors in ``tests/golden/`` that were generated from cv2 by ``tests/golden/gen_match_golden.py``.
"""
from __future__ import annotations

import numpy as np

INT_INF = np.iinfo(np.int32).max


def sqdist_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Exact integer squared L2 distances, int64 [nq, nt].

    fp32 BLAS is exact here: every partial sum of u8*u8 products over 128 dims is an
    integer < 2^24 (128*255^2 = 8 323 200)."""
    q = np.ascontiguousarray(q)
    t = np.ascontiguousarray(t)
    assert q.ndim == 2 and t.ndim == 2 and q.shape[1] == t.shape[1]
    assert q.shape[1] * 255 * 255 < (1 << 24)
    qf = q.astype(np.float32)
    tf = t.astype(np.float32)
    dot = (qf @ tf.T).astype(np.int64)
    nq = (q.astype(np.int64) ** 2).sum(1)
    nt = (t.astype(np.int64) ** 2).sum(1)
    return nq[:, None] + nt[None, :] - 2 * dot


def knn2(q: np.ndarray, t: np.ndarray, chunk: int = 1024):
    """2-NN of every row of q in t.

    Returns (idx [nq,2] int32, dist [nq,2] float32, d2 [nq,2] int64).  Missing neighbours
    (nt < 2) have idx -1, dist inf.  Order is (sqrtf(d2), index) lexicographic, which is what
    OpenCV's stable insertion produces."""
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    d2o = np.full((nq, 2), -1, np.int64)
    if nt == 0 or nq == 0:
        return idx, dist, d2o
    for s in range(0, nq, chunk):
        d2 = sqdist_matrix(q[s:s + chunk], t)
        f = np.sqrt(d2.astype(np.float32))            # correctly rounded, like sqrtf
        # stable argsort on the float distance == (distance, index) lexicographic
        k = min(2, nt)
        if nt > 64:
            # candidates: everything with f <= second-smallest f (ties included), then stable sort those
            part = np.partition(f, 1, axis=1)[:, 1]
            for r in range(f.shape[0]):
                cand = np.nonzero(f[r] <= part[r])[0]
                o = cand[np.argsort(f[r, cand], kind="stable")][:k]
                idx[s + r, :k] = o
                dist[s + r, :k] = f[r, o]
                d2o[s + r, :k] = d2[r, o]
        else:
            o = np.argsort(f, axis=1, kind="stable")[:, :k]
            rows = np.arange(f.shape[0])[:, None]
            idx[s:s + chunk, :k] = o
            dist[s:s + chunk, :k] = f[rows, o]
            d2o[s:s + chunk, :k] = d2[rows, o]
    return idx, dist, d2o


def compute_matches(desc1, desc2, distance_ratio=0.8):
    """FeatureUtils::ComputeMatches (FeatureUtils.cpp:141-157).

    Returns int32 [m,2] (queryIdx, trainIdx) ascending queryIdx and float32 [m] distances.
    ``distance_ratio`` is narrowed to float like the reference's ``const float`` parameter and
    the test is float*float (``m[0].distance < distance_ratio * m[1].distance``, :152).
    With fewer than 2 train rows the reference indexes m[1] out of bounds (UB); defined here as
    "no match"."""
    idx, dist, _ = knn2(desc1, desc2)
    r = np.float32(distance_ratio)
    ok = (idx[:, 1] >= 0) & (dist[:, 0] < (r * dist[:, 1]).astype(np.float32))
    qi = np.nonzero(ok)[0].astype(np.int32)
    return np.stack([qi, idx[qi, 0]], 1).astype(np.int32).reshape(-1, 2), dist[qi, 0]


def cross_check(m12, d12, m21, opencv_quirks=True):
    """FeatureUtils::CrossCheck (FeatureUtils.cpp:281-310).

    ``vis`` is an unordered_map<int,int>; ``vis[train_idx]`` default-inserts 0, so an m12 with
    queryIdx == 0 whose trainIdx has no entry in matches21 is KEPT (opencv_quirks=True
    reproduces that; False gives the mathematically intended mutual check)."""
    vis = {}
    for qi, ti in m21:
        vis[int(qi)] = int(ti)          # later duplicates overwrite, as in the reference (:288-293)
    keep = []
    for k, (qi, ti) in enumerate(m12):
        qi, ti = int(qi), int(ti)
        if opencv_quirks:
            v = vis.setdefault(ti, 0)   # operator[] semantics (:302)
        else:
            v = vis.get(ti, -1)
        if v == qi:
            keep.append(k)
    keep = np.asarray(keep, np.int64)
    return m12[keep].reshape(-1, 2), d12[keep]


def compute_cross_matches(desc1, desc2, distance_ratio=0.8, opencv_quirks=True):
    """FeatureUtils::ComputeCrossMatches (FeatureUtils.cpp:160-174)."""
    m12, d12 = compute_matches(desc1, desc2, distance_ratio)
    m21, _ = compute_matches(desc2, desc1, distance_ratio)
    return cross_check(m12, d12, m21, opencv_quirks)


def filter_matches_by_distance(m, d, max_distance):
    """FeatureUtils::FilterMatchesByDistance (FeatureUtils.cpp:208-218): drop
    ``(double)distance > max_distance``.  max_distance < 0 disables the filter (extension used by the
    u8 contract)."""
    if max_distance is None or max_distance < 0:
        return m, d
    keep = ~(d.astype(np.float64) > float(max_distance))
    return m[keep].reshape(-1, 2), d[keep]


def match_image_pair(desc1, desc2, distance_ratio=0.8, max_distance=-1.0, cross_check_on=True,
                     opencv_quirks=True):
    """Descriptor part of FeatureMatcher::MatchImagePairs (FeatureMatching.cpp:36-49)."""
    if cross_check_on:
        m, d = compute_cross_matches(desc1, desc2, distance_ratio, opencv_quirks)
    else:
        m, d = compute_matches(desc1, desc2, distance_ratio)
    return filter_matches_by_distance(m, d, max_distance)


# ------------------------------------------------------------------------------------------------
# The real thing (OpenCV through its Python binding) — used to pin the restatement and, in
# bench.py, as the CPU baseline.  cv2 is part of this image (also on the GPU box).
# ------------------------------------------------------------------------------------------------
def cv2_knn2(q, t):
    import cv2
    m = cv2.DescriptorMatcher_create("BruteForce")      # FeatureUtils.cpp:146
    res = m.knnMatch(np.ascontiguousarray(q), np.ascontiguousarray(t), 2)   # :149
    nq = q.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    for i, lst in enumerate(res):
        for k, dm in enumerate(lst[:2]):
            idx[i, k] = dm.trainIdx
            dist[i, k] = dm.distance
    return idx, dist


def cv2_compute_matches(desc1, desc2, distance_ratio=0.8):
    idx, dist = cv2_knn2(desc1, desc2)
    r = np.float32(distance_ratio)
    ok = (idx[:, 1] >= 0) & (dist[:, 0] < (r * dist[:, 1]).astype(np.float32))
    qi = np.nonzero(ok)[0].astype(np.int32)
    return np.stack([qi, idx[qi, 0]], 1).astype(np.int32).reshape(-1, 2), dist[qi, 0]


def cv2_match_image_pair(desc1, desc2, distance_ratio=0.8, max_distance=-1.0, cross_check_on=True,
                         opencv_quirks=True):
    m12, d12 = cv2_compute_matches(desc1, desc2, distance_ratio)
    if cross_check_on:
        m21, _ = cv2_compute_matches(desc2, desc1, distance_ratio)
        m12, d12 = cross_check(m12, d12, m21, opencv_quirks)
    return filter_matches_by_distance(m12, d12, max_distance)


# --------------------------------------------------------------------------------------------- extraction-time normalisation
def normalize_descriptors(desc, kind):
    """FeatureExtraction.cpp:260-281 restated with OpenCV's arithmetic (pinned against cv2 in tests/test_oracle_match.py):
    L1RootNormalized: row /= cv::norm(row, NORM_L1); cv::sqrt(row)      L2Normalized: row /= cv::norm(row, NORM_L2).
    cv::norm accumulates in double; `Mat /= double` is convertTo(alpha = 1 / norm), which multiplies 32F data by float(alpha)
    in float."""
    desc = np.asarray(desc, np.float32)
    d64 = desc.astype(np.float64)
    norm = np.abs(d64).sum(1) if kind == "l1_root" else np.sqrt((d64 * d64).sum(1))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = desc * (1.0 / norm).astype(np.float32)[:, None]
        if kind == "l1_root":
            out = np.sqrt(out)
    return out.astype(np.float32)


def cv2_normalize_descriptors(desc, kind):
    """The same through cv2 itself (cv2.normalize = convertTo with alpha / norm, then cv2.sqrt), row by row as the reference."""
    import cv2
    desc = np.asarray(desc, np.float32)
    out = np.empty_like(desc)
    for i in range(len(desc)):
        row = cv2.normalize(desc[i:i + 1], None, 1.0, 0.0, cv2.NORM_L1 if kind == "l1_root" else cv2.NORM_L2)
        out[i] = cv2.sqrt(row) if kind == "l1_root" else row
    return out


def quantize_descriptors(desc):
    """The x512 bridge of the device path: clamp(rint(512 v), 0, 255), round half to even, NaN -> 0."""
    q = np.rint(np.asarray(desc, np.float32) * np.float32(512.0))
    q = np.where(np.isnan(q), 0.0, q)
    return np.clip(q, 0, 255).astype(np.uint8)
