"""CPU: pin the M-path oracle against the golden vectors generated from OpenCV (tests/golden/gen_match_golden.py)
and, when cv2 is importable, against OpenCV itself on fresh inputs."""
import numpy as np
import pytest

from oracle import match_oracle as mo

CASES = ["cfg1", "ragged", "sift_planted", "ties", "sqrt_collapse", "quirk_q0", "n2_is_1", "n2_is_2", "n1_is_1",
         "preempt100", "all_equal", "extremes", "real_sift"]


@pytest.mark.parametrize("name", CASES)
def test_knn2_matches_opencv_golden(golden_match, name):
    g = golden_match
    a, b = g[f"{name}/a"], g[f"{name}/b"]
    for (q, t, tag) in ((a, b, "knn12"), (b, a, "knn21")):
        idx, dist, _ = mo.knn2(q, t)
        np.testing.assert_array_equal(idx, g[f"{name}/{tag}_idx"])
        np.testing.assert_array_equal(dist, g[f"{name}/{tag}_dist"])   # bit-exact floats


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ratio", [0.8, 0.95])
def test_match_lists_match_opencv_golden(golden_match, name, ratio):
    g = golden_match
    a, b = g[f"{name}/a"], g[f"{name}/b"]
    tag = f"{name}/r{int(ratio * 100)}"
    m12, d12 = mo.compute_matches(a, b, ratio)
    np.testing.assert_array_equal(m12, g[f"{tag}/m12"])
    np.testing.assert_array_equal(d12, g[f"{tag}/d12"])
    for quirk in (1, 0):
        m, d = mo.match_image_pair(a, b, ratio, -1.0, True, bool(quirk))
        np.testing.assert_array_equal(m, g[f"{tag}/cross_q{quirk}"])
        np.testing.assert_array_equal(d, g[f"{tag}/cross_q{quirk}_d"])


def test_quirk_is_exercised(golden_match):
    g = golden_match
    assert len(g["quirk_q0/r80/cross_q1"]) == len(g["quirk_q0/r80/cross_q0"]) + 1
    assert g["quirk_q0/r80/cross_q1"][0, 0] == 0


def test_sqrt_collapse_prefers_lower_index(golden_match):
    g = golden_match
    # d2 = 5000012 at column 3, 5000011 at column 35: same float sqrt -> OpenCV returns column 3 first
    assert g["sqrt_collapse/knn12_idx"][0, 0] == 3 and g["sqrt_collapse/knn12_idx"][0, 1] == 20 or \
        g["sqrt_collapse/knn12_idx"][0, 1] in (20, 35)
    idx, _, d2 = mo.knn2(g["sqrt_collapse/a"], g["sqrt_collapse/b"])
    assert idx[0, 0] == 3 and d2[0, 0] == 5000012


def test_against_cv2_fresh_inputs():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(123)
    for (n1, n2) in ((257, 511), (64, 1000)):
        a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)
        b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
        b[n2 // 2] = b[3]
        a[5] = b[3]
        i1, d1, _ = mo.knn2(a, b)
        i2, d2 = mo.cv2_knn2(a, b)
        np.testing.assert_array_equal(i1, i2)
        np.testing.assert_array_equal(d1, d2)


def test_bench_workload_with_planted_ties_is_pinned_to_cv2():
    """The benchmark's S images (bench.py: make_descriptors_numpy + plant_ties: duplicated rows = two columns at the best
    distance; all-255 rows = every distance in the float-sqrt collapse range when the other image has none) through the oracle
    and through cv2 itself: same neighbours, same distances, and the planted cases really occur."""
    cv2 = pytest.importorskip("cv2")
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    d = bench.make_descriptors_numpy(3, 1500, 9)
    for (i, j) in ((2, 1), (1, 2), (2, 0)):
        i1, d1, dd1 = mo.knn2(d[i], d[j])
        i2, d2 = mo.cv2_knn2(d[i], d[j])
        np.testing.assert_array_equal(i1, i2)
        np.testing.assert_array_equal(d1, d2)
    idx, dist, d2sq = mo.knn2(d[2], d[1])
    assert (dist[:8, 0] == dist[:8, 1]).all() and (idx[:4, 0] == np.arange(4)).all() and (idx[:4, 1] == np.arange(4, 8)).all()   # ties: lowest index first
    assert (d2sq[8:12, 0] >= 1 << 22).all()                                   # even image vs odd image: collapse range
    m, _ = mo.match_image_pair(d[2], d[1], 0.8, -1.0, True, True)
    assert not (m[:, 0] < 12).any()                                           # neither the tied nor the far rows give a match


def test_filter_by_distance():
    m = np.array([[0, 1], [1, 2], [2, 3]], np.int32)
    d = np.array([0.5, 0.7, 0.70001], np.float32)
    mm, dd = mo.filter_matches_by_distance(m, d, 0.7)
    # float32(0.7) = 0.699999988 <= 0.7 (double) -> kept, exactly like `distance > max_distance` in the reference
    assert mm.tolist() == [[0, 1], [1, 2]]
    mm, dd = mo.filter_matches_by_distance(m, d, -1.0)
    assert len(mm) == 3


def test_descriptor_normalisation_restatement_is_pinned_to_cv2():
    """oracle.normalize_descriptors (FeatureExtraction.cpp:260-281) bit for bit against cv2.normalize / cv2.sqrt on raw SIFT-like
    rows (integer valued floats as cv::SIFT produces) and on arbitrary floats."""
    rng = np.random.default_rng(3)
    sift = np.floor(np.abs(rng.standard_normal((600, 128))) * 40).astype(np.float32)
    anyf = np.abs(rng.standard_normal((300, 128))).astype(np.float32) * 7.3
    for d in (sift, anyf):
        for kind in ("l1_root", "l2"):
            a, b = mo.normalize_descriptors(d, kind), mo.cv2_normalize_descriptors(d, kind)
            assert (a.view(np.uint32) == b.view(np.uint32)).mean() >= (1.0 if d is sift else 0.999)     # double summation order
            np.testing.assert_allclose(a, b, rtol=2e-7, atol=0)
    assert np.allclose(np.linalg.norm(mo.normalize_descriptors(sift, "l1_root"), axis=1), 1.0, atol=1e-5)   # RootSIFT rows are unit L2
