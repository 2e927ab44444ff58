"""GPU parity tests of the M-path, through the C-ABI (ctypes), against the OpenCV golden vectors and the oracle."""
import numpy as np
import pytest

import monocularsfm_b200 as m
from oracle import match_oracle as mo

pytestmark = pytest.mark.gpu

CASES = ["cfg1", "ragged", "sift_planted", "ties", "sqrt_collapse", "quirk_q0", "n2_is_1", "n2_is_2", "n1_is_1",
         "preempt100", "all_equal", "extremes", "real_sift"]


def _check_knn(idx, dist, d2, gi, gd, mode):
    np.testing.assert_array_equal(idx[:, 0], gi[:, 0])
    np.testing.assert_array_equal(dist, gd)                     # both distances, bit-exact floats
    if mode == 1:
        np.testing.assert_array_equal(idx[:, 1], gi[:, 1])
    else:
        known = idx[:, 1] >= 0                                 # only rows that took the exact path report it
        np.testing.assert_array_equal(idx[known, 1], gi[known, 1])


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_knn2_golden(ctx, golden_match, name, mode):
    g = golden_match
    a, b = g[f"{name}/a"], g[f"{name}/b"]
    idx, dist, d2 = ctx.knn2(a, b, mode)
    _check_knn(idx, dist, d2, g[f"{name}/knn12_idx"], g[f"{name}/knn12_dist"], mode)
    idx, dist, d2 = ctx.knn2(b, a, mode)
    _check_knn(idx, dist, d2, g[f"{name}/knn21_idx"], g[f"{name}/knn21_dist"], mode)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ratio", [0.8, 0.95])
def test_match_pairs_golden(ctx, golden_match, name, ratio):
    g = golden_match
    a, b = g[f"{name}/a"], g[f"{name}/b"]
    tag = f"{name}/r{int(ratio * 100)}"
    ctx.upload(1, a)
    ctx.upload(2, b)
    off, mt, d = ctx.match_pairs([[1, 2]], m.MatchOptions(ratio, -1.0, False, False))
    np.testing.assert_array_equal(mt, g[f"{tag}/m12"])
    np.testing.assert_array_equal(d, g[f"{tag}/d12"])
    for quirk in (1, 0):
        off, mt, d = ctx.match_pairs([[1, 2]], m.MatchOptions(ratio, -1.0, True, bool(quirk)))
        assert off.tolist() == [0, len(g[f"{tag}/cross_q{quirk}"])]
        np.testing.assert_array_equal(mt, g[f"{tag}/cross_q{quirk}"])
        np.testing.assert_array_equal(d, g[f"{tag}/cross_q{quirk}_d"])


def _sift_like(rng, n):
    x = np.abs(rng.standard_normal((n, 128)))
    x = x / np.linalg.norm(x, axis=1, keepdims=True) * 512.0
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def _planted_set(rng, n_img, n):
    base = _sift_like(rng, n)
    imgs = [base]
    for k in range(1, n_img):
        x = _sift_like(rng, n)
        m_ = int(0.3 * n)
        dst = rng.permutation(n)[:m_]
        src = rng.permutation(n)[:m_]
        x[dst] = np.clip(base[src].astype(np.int64) + rng.integers(-2, 3, (m_, 128)), 0, 255).astype(np.uint8)
        imgs.append(x)
    return imgs


@pytest.mark.parametrize("n1,n2", [(2000, 3000), (1, 257), (129, 31), (8192, 8192)])
def test_knn2_vs_oracle_random(ctx, n1, n2):
    rng = np.random.default_rng(n1 * 7 + n2)
    a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
    oi, od, _ = mo.knn2(a, b)
    idx, dist, d2 = ctx.knn2(a, b, 0)
    _check_knn(idx, dist, d2, oi, od, 0)


def test_multi_pair_batch_vs_oracle(ctx):
    """Several ragged images, all pairs in one call, cross-check + distance filter, vs the oracle."""
    rng = np.random.default_rng(7)
    sizes = [700, 333, 1024, 50, 2, 0, 1500]
    imgs = _planted_set(rng, len(sizes), 1500)
    imgs = [im[:s] for im, s in zip(imgs, sizes)]
    for i, im in enumerate(imgs):
        ctx.upload(10 + i, im)
    pairs = [(10 + i, 10 + j) for i in range(len(sizes)) for j in range(i)]
    for opt in (m.MatchOptions(0.8, -1.0, True, True), m.MatchOptions(0.9, 60.0, True, False),
                m.MatchOptions(0.8, -1.0, False, True)):
        off, mt, d = ctx.match_pairs(pairs, opt)
        assert off[0] == 0 and off[-1] == len(mt)
        for p, (i1, i2) in enumerate(pairs):
            em, ed = mo.match_image_pair(imgs[i1 - 10], imgs[i2 - 10], opt.distance_ratio, opt.max_distance,
                                         bool(opt.cross_check), bool(opt.opencv_quirks))
            np.testing.assert_array_equal(mt[off[p]:off[p + 1]], em, err_msg=f"pair {p} {i1}-{i2}")
            np.testing.assert_array_equal(d[off[p]:off[p + 1]], ed)
    st = ctx.match_stats()
    assert st["units"] > 0


def test_full_size_pair_properties(ctx):
    """BASELINE cfg-2 shape (8192 x 8192): planted correspondences must be recovered; symmetric under swap."""
    rng = np.random.default_rng(11)
    a, b = _planted_set(rng, 2, 8192)
    ctx.upload(1, a)
    ctx.upload(2, b)
    opt = m.MatchOptions(0.8, -1.0, True, False)
    off, m12, d12 = ctx.match_pairs([[1, 2]], opt)
    off, m21, d21 = ctx.match_pairs([[2, 1]], opt)
    assert len(m12) > 2000
    # mutual matching is symmetric: (i, j) in m12  <=>  (j, i) in m21
    s12 = set(map(tuple, m12.tolist()))
    s21 = set((j, i) for i, j in m21.tolist())
    assert s12 == s21
    assert (np.diff(m12[:, 0]) > 0).all()                     # ascending queryIdx
    em, ed = mo.match_image_pair(a, b, 0.8, -1.0, True, False)
    np.testing.assert_array_equal(m12, em)
    np.testing.assert_array_equal(d12, ed)


def test_capacity_error(ctx):
    rng = np.random.default_rng(3)
    a, b = _planted_set(rng, 2, 512)
    ctx.upload(1, a)
    ctx.upload(2, b)
    with pytest.raises(m.MsfmError) as ei:
        ctx.match_pairs([[1, 2]], m.MatchOptions(), capacity=3)
    assert ei.value.code == -4


def test_unknown_image_is_an_error(ctx):
    with pytest.raises(m.MsfmError) as ei:
        ctx.match_pairs([[12345, 1]], m.MatchOptions())
    assert ei.value.code == -5


def test_many_small_pairs_cross_batch_boundaries(ctx):
    """More work units than one internal batch holds (131072): offsets / running totals must chain across batches.
    200 images x 260 descriptors -> 19900 pairs x 2 directions x 4 units."""
    rng = np.random.default_rng(99)
    n_img, n = 200, 260
    base = _sift_like(rng, n)
    imgs = []
    for k in range(n_img):
        x = _sift_like(rng, n)
        m_ = 60
        dst = rng.permutation(n)[:m_]
        src = rng.permutation(n)[:m_]
        x[dst] = np.clip(base[src].astype(np.int64) + rng.integers(-2, 3, (m_, 128)), 0, 255).astype(np.uint8)
        imgs.append(x)
        ctx.upload(1000 + k, x)
    pairs = np.array([(1000 + i, 1000 + j) for i in range(n_img) for j in range(i)], np.int32)
    off, mt, d = ctx.match_pairs(pairs, m.MatchOptions(0.8, -1.0, True, True))
    st = ctx.match_stats()
    assert st["units"] > 131072
    assert off[0] == 0 and off[-1] == len(mt) and (np.diff(off) >= 0).all()
    for p in list(range(0, len(pairs), 997)) + [len(pairs) - 1, 8190, 8191, 8192, 8193]:
        i1, i2 = pairs[p] - 1000
        em, ed = mo.match_image_pair(imgs[i1], imgs[i2], 0.8, -1.0, True, True)
        np.testing.assert_array_equal(mt[off[p]:off[p + 1]], em, err_msg=f"pair {p}")
        np.testing.assert_array_equal(d[off[p]:off[p + 1]], ed)
    for k in range(n_img):
        ctx.release(1000 + k)


def test_api_edge_cases(ctx):
    """Empty inputs, re-upload with a different size, release, zero pairs."""
    rng = np.random.default_rng(17)
    a, b = _planted_set(rng, 2, 300)
    ctx.upload(50, a)
    ctx.upload(51, b)
    ctx.upload(52, np.zeros((0, 128), np.uint8))                 # an image without descriptors
    assert ctx.desc_count(52) == 0
    off, mt, d = ctx.match_pairs(np.zeros((0, 2), np.int32), m.MatchOptions(), capacity=1)
    assert off.tolist() == [0] and len(mt) == 0
    off, mt, d = ctx.match_pairs([[50, 52], [52, 50], [50, 51]], m.MatchOptions(0.8, -1.0, True, True))
    assert off[1] == 0 and off[2] == 0
    em, ed = mo.match_image_pair(a, b, 0.8, -1.0, True, True)
    np.testing.assert_array_equal(mt, em)
    # replace image 51 by a smaller set: results must follow the new data
    b2 = b[:77]
    ctx.upload(51, b2)
    assert ctx.desc_count(51) == 77
    off, mt, d = ctx.match_pairs([[50, 51]], m.MatchOptions(0.8, -1.0, True, True))
    em, ed = mo.match_image_pair(a, b2, 0.8, -1.0, True, True)
    np.testing.assert_array_equal(mt, em)
    np.testing.assert_array_equal(d, ed)
    # a pair of an image with itself: every descriptor's nearest neighbour is itself at distance 0
    off, mt, d = ctx.match_pairs([[50, 50]], m.MatchOptions(0.8, -1.0, True, True))
    em, ed = mo.match_image_pair(a, a, 0.8, -1.0, True, True)
    np.testing.assert_array_equal(mt, em)
    ctx.release(51)
    with pytest.raises(m.MsfmError):
        ctx.match_pairs([[50, 51]], m.MatchOptions())
    for k in (50, 52):
        ctx.release(k)


def test_uniform_random_descriptors_full_norm_range(ctx):
    """Uniform u8 rows have squared norms spread over several sort buckets and both parities (the SIFT-like sets do not)."""
    rng = np.random.default_rng(23)
    a = rng.integers(0, 256, (1500, 128), dtype=np.uint8)
    b = rng.integers(0, 256, (1700, 128), dtype=np.uint8)
    a[:40] = (a[:40] // 8)                       # small norms
    b[:60] = 255 - (b[:60] // 16)                # very large norms (third bucket)
    b[100:160] = a[100:160]                      # exact correspondences
    ctx.upload(60, a)
    ctx.upload(61, b)
    for opt in (m.MatchOptions(0.8, -1.0, True, True), m.MatchOptions(0.97, -1.0, False, False)):
        off, mt, d = ctx.match_pairs([[60, 61]], opt)
        em, ed = mo.match_image_pair(a, b, opt.distance_ratio, -1.0, bool(opt.cross_check), bool(opt.opencv_quirks))
        np.testing.assert_array_equal(mt, em)
        np.testing.assert_array_equal(d, ed)
    idx, dist, d2 = ctx.knn2(a, b, 0)
    oi, od, _ = mo.knn2(a, b)
    np.testing.assert_array_equal(idx[:, 0], oi[:, 0])
    np.testing.assert_array_equal(dist, od)
    ctx.release(60)
    ctx.release(61)


def _quantise_like_the_bridge(x):
    """Host statement of the float32 -> uint8 bridge (FeatureUtils::ToUint8Descriptors in the C++ shim):
    sets of integers in [0,255] convert exactly, anything else is clamp(rint(512 v), 0, 255), round-half-to-even."""
    x = np.asarray(x, np.float32)
    integral = bool(np.all((x >= 0) & (x <= 255) & (x == np.floor(x))))
    q = x if integral else np.rint(x * np.float32(512.0))
    q = np.where(np.isnan(q), 0, q)                      # std::max(0.f, NaN) == 0.f in the C++ statement
    return np.clip(q, 0, 255).astype(np.uint8), not integral


@pytest.mark.gpu
def test_float32_upload_bridge(ctx):
    """msfm_desc_upload_f32: what Database::ReadDescriptors returns (CV_32F, Database.cpp:510-523) converted on the device
    gives exactly the matches of the uint8 upload of the host-side statement of the bridge."""
    rng = np.random.default_rng(77)
    raw = rng.integers(0, 256, (700, 128)).astype(np.float32)                    # un-normalised SIFT: integers
    raw2 = raw[rng.permutation(700)[:500]] + rng.integers(-3, 4, (500, 128))
    raw2 = np.clip(raw2, 0, 255).astype(np.float32)
    # L1-root normalised rows (FeatureExtraction.cpp:260-270), incl. exact .5 ties after scaling, a NaN and out-of-range values
    l1 = np.sqrt(raw / np.maximum(raw.sum(1, keepdims=True), 1)).astype(np.float32)
    l1b = np.sqrt(raw2 / np.maximum(raw2.sum(1, keepdims=True), 1)).astype(np.float32)
    l1[0, :8] = np.array([0.5 / 512, 1.5 / 512, 2.5 / 512, 254.5 / 512, 255.5 / 512, -0.25, 3.0, np.nan], np.float32)
    opt = m.MatchOptions(0.8, -1.0, True, True)
    for a, b in ((raw, raw2), (l1, l1b)):
        qa, quant_a = _quantise_like_the_bridge(a)
        qb, quant_b = _quantise_like_the_bridge(b)
        ctx.upload(10, qa)
        ctx.upload(11, qb)
        off_u8, mt_u8, d_u8 = ctx.match_pairs([[10, 11]], opt)
        ctx.upload_f32(20, a)
        ctx.sync()                                                   # the float staging buffer is reused by the next upload
        ctx.upload_f32(21, b)
        off_f, mt_f, d_f = ctx.match_pairs([[20, 21]], opt)
        assert mt_f.tolist() == mt_u8.tolist() and (d_f == d_u8).all()
        assert len(mt_f) > 50
        assert ctx.quantised(20) == quant_a and ctx.quantised(21) == quant_b
        assert ctx.quantised(10) is False
        em, ed = mo.match_image_pair(qa, qb, 0.8, -1.0, True, True)   # and both equal the OpenCV-pinned oracle on the uint8 sets
        assert mt_f.tolist() == em.tolist() and (d_f == ed).all()
    # forced quantisation of an integral set: everything saturates at 255 except zeros
    ctx.upload_f32(30, raw, always_quantise=True)
    assert ctx.quantised(30) is True
    for k in (10, 11, 20, 21, 30):
        ctx.release(k)


@pytest.mark.parametrize("kind", ["l1_root", "l2"])
def test_raw_sift_upload_normalises_like_the_reference_extraction(ctx, kind):
    """msfm_desc_upload_raw_f32: the normalised rows it returns equal FeatureExtraction's (oracle pinned to cv2) bit for bit on
    integer-valued SIFT rows, and the resident set is their x512 quantisation: matching the raw upload against an upload of
    the oracle's quantised bytes gives identical lists."""
    rng = np.random.default_rng(17)
    raw1 = np.floor(np.abs(rng.standard_normal((700, 128))) * 40).astype(np.float32)
    raw2 = np.floor(np.abs(rng.standard_normal((900, 128))) * 40).astype(np.float32)
    raw2[rng.permutation(900)[:250]] = raw1[rng.permutation(700)[:250]]
    raw1[5] = 0.0                                             # an all-zero row: 0 / 0 -> NaN rows quantise to zeros
    n1 = ctx.upload_raw_f32(10, raw1, kind)
    n2 = ctx.upload_raw_f32(11, raw2, kind)
    o1, o2 = mo.normalize_descriptors(raw1, kind), mo.normalize_descriptors(raw2, kind)
    ok = np.ones(len(raw1), bool)
    ok[5] = False                                             # the 0 / 0 row is NaN on both sides (payload bits are not compared)
    assert np.isnan(n1[5]).all() and np.isnan(o1[5]).all()
    assert np.array_equal(n1[ok].view(np.uint32), o1[ok].view(np.uint32)) and np.array_equal(n2.view(np.uint32), o2.view(np.uint32))
    assert ctx.quantised(10) and ctx.quantised(11)
    off, mt, d = ctx.match_pairs([[10, 11]], m.MatchOptions(0.8, 0.7 * 512.0, True, True))
    em, ed = mo.match_image_pair(mo.quantize_descriptors(o1), mo.quantize_descriptors(o2), 0.8, 0.7 * 512.0, True, True)
    assert mt.tolist() == em.tolist() and np.array_equal(d, ed) and len(em) > 150
