#!/usr/bin/env python
"""Generate the M-path golden vectors FROM OpenCV ITSELF (cv2.BFMatcher, the routine the
reference calls at src/Feature/FeatureUtils.cpp:146-149).  Run in the dev container:

    python tests/golden/gen_match_golden.py

Writes tests/golden/match_golden.npz.  The fixtures pin oracle/match_oracle.py and the CUDA
path; they are what "parity" means for the M-path (the reference's own tests hold no vectors
for it).  Inputs are stored next to outputs so nothing depends on RNG
reproducibility across numpy versions.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import match_oracle as mo  # noqa: E402  (only for the restated ratio/cross-check glue)


def vec_with_sqnorm(d2):
    """u8 vector (128-D) whose squared norm is exactly d2 (greedy sum of squares)."""
    v = np.zeros(128, np.int64)
    rem = int(d2)
    for k in range(128):
        x = int(min(255, np.floor(np.sqrt(rem))))
        v[k] = x
        rem -= x * x
        if rem == 0:
            break
    assert rem == 0, (d2, rem)
    return v.astype(np.uint8)


def sift_like(rng, n):
    x = np.abs(rng.standard_normal((n, 128)))
    x = x / np.linalg.norm(x, axis=1, keepdims=True) * 512.0
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def real_sift_pair(seed=5):
    """Two synthetic textured views (random blobs + noise; the second one an affine warp of the first with fresh noise)
    through cv2.SIFT — the detector/descriptor FeatureUtils::ExtractFeature uses (src/Feature/FeatureUtils.cpp:14-36).
    Real SIFT output is integer-valued float32 in [0, 255]; stored as uint8 (lossless)."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.zeros((480, 640), np.float32)
    for _ in range(400):
        c = (int(rng.integers(0, 640)), int(rng.integers(0, 480)))
        cv2.circle(img, c, int(rng.integers(3, 25)), float(rng.uniform(40, 255)), -1)
        p1 = (int(rng.integers(0, 640)), int(rng.integers(0, 480)))
        p2 = (int(rng.integers(0, 640)), int(rng.integers(0, 480)))
        cv2.line(img, p1, p2, float(rng.uniform(0, 255)), int(rng.integers(1, 4)))
    img = cv2.GaussianBlur(img, (0, 0), 1.2)
    a_img = np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
    M = cv2.getRotationMatrix2D((320, 240), 7.0, 1.08)
    M[:, 2] += (9.0, -6.0)
    b_img = cv2.warpAffine(img, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    b_img = np.clip(b_img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
    sift = cv2.SIFT_create()
    _, da = sift.detectAndCompute(a_img, None)
    _, db = sift.detectAndCompute(b_img, None)
    assert da is not None and db is not None and (da == np.floor(da)).all() and da.max() <= 255 and da.min() >= 0
    return da.astype(np.uint8), db.astype(np.uint8)


def cases():
    rng = np.random.default_rng(0)
    out = {}
    # cfg1 of BASELINE.json: 2 images x 512 x 128 u8, seed 0, uniform
    a = rng.integers(0, 256, (512, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (512, 128), dtype=np.uint8)
    out["cfg1"] = (a, b)
    # ragged sizes, not multiples of the 128/256 tiles
    out["ragged"] = (rng.integers(0, 256, (300, 128), dtype=np.uint8),
                     rng.integers(0, 256, (777, 128), dtype=np.uint8))
    # SIFT-like with planted correspondences + noise (non-trivial ratio / cross-check lists)
    s0 = sift_like(rng, 600)
    s1 = sift_like(rng, 700)
    perm = rng.permutation(700)[:200]
    src = rng.permutation(600)[:200]
    s1[perm] = np.clip(s0[src].astype(np.int64) + rng.integers(-2, 3, (200, 128)), 0, 255).astype(np.uint8)
    out["sift_planted"] = (s0, s1)
    # exact duplicates in the train set (tie -> lowest index), duplicates spanning tile borders
    a = sift_like(rng, 400)
    b = sift_like(rng, 600)
    b[5] = a[0]; b[300] = a[0]; b[599] = a[0]          # three-way exact tie for query 0
    b[40] = a[1]; b[41] = a[1]                          # adjacent tie
    b[255] = a[2]; b[256] = a[2]                        # tie across a 256-column tile border
    b[31] = a[3]; b[32] = a[3]                          # tie across a 32-column group border
    a[399] = a[0]
    out["ties"] = (a, b)
    # sqrt collapse: d2 >= 2^22 -> distinct integers share one float sqrt; lower index wins
    q = np.zeros((4, 128), np.uint8)
    t = np.zeros((40, 128), np.uint8)
    for j in range(40):
        t[j] = vec_with_sqnorm(6000000 + 1000 * j)
    t[3] = vec_with_sqnorm(5000012)
    t[35] = vec_with_sqnorm(5000011)                    # smaller integer, larger index, same sqrtf
    t[20] = vec_with_sqnorm(5000013)
    out["sqrt_collapse"] = (q, t)
    # queryIdx==0 cross-check quirk: query 0 has a ratio-passing match whose reverse fails
    a = sift_like(rng, 64)
    b = sift_like(rng, 64)
    a[0] = np.minimum(a[0], 250)
    b[10] = a[0]; b[10, :64] += 1                       # 0 -> 10 passes (d = 8, runner-up far away)
    a[50] = a[0]; a[50, :64] += 2                       # reverse: 10 -> {0, 50} both at d = 8 -> ratio fails,
                                                        # so trainIdx 10 has NO entry in matches21 -> vis[10] == 0 == queryIdx
    out["quirk_q0"] = (a, b)
    # tiny train sets
    out["n2_is_1"] = (sift_like(rng, 5), sift_like(rng, 1))
    out["n2_is_2"] = (sift_like(rng, 5), sift_like(rng, 2))
    out["n1_is_1"] = (sift_like(rng, 1), sift_like(rng, 33))
    # preemptive-matching shape (100 x 100, FeatureMatching.h:98)
    out["preempt100"] = (sift_like(rng, 100), sift_like(rng, 100))
    # all-equal descriptors: every column ties
    out["all_equal"] = (np.full((40, 128), 7, np.uint8), np.full((300, 128), 7, np.uint8))
    # extremes 0 / 255 (max d2 = 8 323 200)
    a = np.zeros((8, 128), np.uint8); b = np.full((9, 128), 255, np.uint8); b[4, :64] = 0
    out["extremes"] = (a, b)
    # real SIFT descriptors of two views of a synthetic scene (true correspondences, repeated structure, many zeros)
    out["real_sift"] = real_sift_pair()
    return out


def main():
    import cv2
    store = {"cv2_version": np.array(cv2.__version__)}
    for name, (a, b) in cases().items():
        idx12, dist12 = mo.cv2_knn2(a, b)
        idx21, dist21 = mo.cv2_knn2(b, a)
        store[f"{name}/a"] = a
        store[f"{name}/b"] = b
        store[f"{name}/knn12_idx"] = idx12
        store[f"{name}/knn12_dist"] = dist12
        store[f"{name}/knn21_idx"] = idx21
        store[f"{name}/knn21_dist"] = dist21
        for ratio in (0.8, 0.95):
            tag = f"{name}/r{int(ratio * 100)}"
            m12, d12 = mo.cv2_compute_matches(a, b, ratio)
            store[f"{tag}/m12"] = m12
            store[f"{tag}/d12"] = d12
            for quirk in (1, 0):
                m, d = mo.cv2_match_image_pair(a, b, ratio, -1.0, True, bool(quirk))
                store[f"{tag}/cross_q{quirk}"] = m
                store[f"{tag}/cross_q{quirk}_d"] = d
        print(name, a.shape, b.shape, "m12@0.8:", len(store[f"{name}/r80/m12"]),
              "cross:", len(store[f"{name}/r80/cross_q1"]), len(store[f"{name}/r80/cross_q0"]))
    path = os.path.join(os.path.dirname(__file__), "match_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
