"""The C++ classes that keep the reference's names (monocularsfm_b200/host), driven through build/host_test.
CPU: Database format / pair ids / CrossCheck quirk.  GPU: FeatureUtils, Brute/SequentialFeatureMatcher on a database in the
reference's schema (written here with Python's sqlite3), CeresBundelOptimizer::Optimize on a BundleData."""
import os
import sqlite3
import struct
import subprocess

import numpy as np
import pytest

from oracle import ba_oracle as bo
from oracle import match_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "host_test")


def _need_exe():
    if not os.path.exists(EXE):
        pytest.skip("build/host_test not built (make)")


def test_host_cpu(tmp_path):
    _need_exe()
    db = str(tmp_path / "t.db")
    out = subprocess.run([EXE, "cpu", db], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    # the file is a plain SQLite database in the reference's schema (Database.cpp:710-764)
    con = sqlite3.connect(db)
    names = {r[0] for r in con.execute("select name from sqlite_master where type='table'")}
    assert {"images", "keypoints", "colors", "descriptors", "matches"} <= names
    rows, cols, data = con.execute("select rows, cols, data from descriptors where image_id = 1").fetchone()
    d = np.frombuffer(data, np.float32).reshape(rows, cols)
    assert (rows, cols) == (3, 128) and d[2, 5] == 19.0
    pid, r, c, blob = con.execute("select pair_id, rows, cols, data from matches").fetchone()
    assert pid == 10002 and (r, c) == (2, 2)
    assert np.frombuffer(blob, np.int32).reshape(2, 2).tolist() == [[7, 5], [1, 6]]     # stored with id1 < id2 orientation


def _sift_like(rng, n):
    x = np.abs(rng.standard_normal((n, 128)))
    x = x / np.linalg.norm(x, axis=1, keepdims=True) * 512.0
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def _make_db(path, rng, sizes, normalised=False):
    """Reference-format database: image ids 0..N-1 (FeatureExtraction writes explicit ids), float32 descriptor blobs —
    integral values (un-normalised SIFT) or, with normalised=True, the same rows divided by 512 (unit-norm floats as the
    reference's extraction stores them; the x512 quantisation of the device bridge recovers the integers exactly) —
    and keypoints with distinct sizes."""
    con = sqlite3.connect(path)
    con.executescript("""
        CREATE TABLE images(image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE);
        CREATE TABLE keypoints(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE colors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE descriptors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE matches(pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
    """)
    base = _sift_like(rng, max(sizes))
    descs, scales = [], []
    for i, n in enumerate(sizes):
        d = _sift_like(rng, n)
        m = int(0.4 * min(n, len(base)))
        dst = rng.permutation(n)[:m]
        src = rng.permutation(len(base))[:m]
        src[:60] = np.arange(60)                                             # every image shares base rows 0..59 ...
        d[dst] = np.clip(base[src].astype(np.int64) + rng.integers(-2, 3, (m, 128)), 0, 255).astype(np.uint8)
        kp = np.zeros((n, 4), np.float32)
        kp[:, :2] = rng.uniform(0, 1000, (n, 2))
        kp[:, 2] = rng.permutation(n).astype(np.float32) + 1.0            # distinct scales
        if i != len(sizes) - 1:
            kp[dst[:60], 2] += 10000.0                                       # ... among its largest-scale features,
        else:
            kp[dst, 2] *= 1e-4                                               # except the last image, whose planted rows
                                                                             # are its smallest: preemption drops its pairs
        con.execute("insert into images(image_id, name) values(?, ?)", (i, f"img{i}.jpg"))
        con.execute("insert into keypoints values(?,?,?,?)", (i, n, 4, kp.tobytes()))
        con.execute("insert into descriptors values(?,?,?,?)",
                    (i, n, 128, (d.astype(np.float32) / (512.0 if normalised else 1.0)).astype(np.float32).tobytes()))
        descs.append(d)
        scales.append(kp[:, 2])
    con.commit()
    con.close()
    return descs, scales


def _read_matches(path):
    con = sqlite3.connect(path)
    out = {}
    for pid, r, c, blob in con.execute("select pair_id, rows, cols, data from matches"):
        out[pid] = np.frombuffer(blob or b"", np.int32).reshape(r, 2)
    con.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("preempt", [0, 1])
def test_brute_feature_matcher_on_reference_database(tmp_path, preempt):
    _need_exe()
    rng = np.random.default_rng(21 + preempt)
    db = str(tmp_path / "m.db")
    sizes = [900, 700, 1100, 300]
    descs, scales = _make_db(db, rng, sizes, normalised=True)
    out = subprocess.run([EXE, "match", db, str(preempt), "noverify"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    got = _read_matches(db)
    n_kept = 0
    for i in range(len(sizes)):
        for j in range(i):
            pid = 10000 * j + i
            keep = True
            if preempt:
                # FeatureMatching.cpp:148-179: 100 largest-scale descriptors each, cross matching, >= 4 matches
                t1 = descs[i][np.argsort(-scales[i], kind="stable")[:100]]
                t2 = descs[j][np.argsort(-scales[j], kind="stable")[:100]]
                keep = len(mo.match_image_pair(t1, t2, 0.8, -1.0, True, True)[0]) >= 4
            if not keep:
                assert pid not in got
                continue
            n_kept += 1
            # MatchImagePairs(i, j): query = image i, train = image j; max_distance 0.7 of the unit-norm floats = 0.7 x 512 on
            # the quantised scale
            em, _ = mo.match_image_pair(descs[i], descs[j], 0.8, 0.7 * 512.0, True, True)
            stored = em[:, ::-1] if i > j else em                             # swapped to id1 < id2 orientation on disk
            np.testing.assert_array_equal(got[pid], stored, err_msg=f"pair {i}-{j}")
    assert n_kept == (3 if preempt else 6), n_kept        # preemption keeps the pairs among images 0..2 only
    # resume: a second run finds every row and changes nothing
    out = subprocess.run([EXE, "match", db, str(preempt), "noverify"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "Existing" in out.stdout
    again = _read_matches(db)
    assert set(again) == set(got)


@pytest.mark.gpu
def test_matcher_verifies_geometry_by_default(tmp_path):
    """Without any call to SetGeometricFilter the matcher runs the geometric verification (FeatureMatching.cpp:60) — batched
    on the device — before WriteMatches: descriptor matches whose keypoints do not fit the two-view geometry are NOT stored."""
    _need_exe()
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_verify_gpu import two_view
    rng = np.random.default_rng(77)
    n_in, n_out = 300, 200
    kp1, kp2, matches, is_in = two_view(rng, n_in, n_out)
    n = n_in + n_out
    d1 = _sift_like(rng, n)
    d2 = _sift_like(rng, n)
    d2[matches[:, 1]] = d1[matches[:, 0]]                       # every (true or wrong) correspondence is a perfect descriptor match
    db = str(tmp_path / "v.db")
    con = sqlite3.connect(db)
    con.executescript("""
        CREATE TABLE images(image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE);
        CREATE TABLE keypoints(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE colors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE descriptors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE matches(pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
    """)
    for i, (kp, d) in enumerate(((kp1, d1), (kp2, d2))):
        k4 = np.zeros((n, 4), np.float32)
        k4[:, :2] = kp
        k4[:, 2] = 2.0
        con.execute("insert into images(image_id, name) values(?, ?)", (i, f"img{i}.jpg"))
        con.execute("insert into keypoints values(?,?,?,?)", (i, n, 4, k4.tobytes()))
        con.execute("insert into descriptors values(?,?,?,?)", (i, n, 128, (d.astype(np.float32) / 512.0).astype(np.float32).tobytes()))
    con.commit()
    con.close()
    out = subprocess.run([EXE, "match", db, "0"], capture_output=True, text=True, timeout=300)      # default: verification on
    assert out.returncode == 0, out.stderr + out.stdout
    got = _read_matches(db)[1]                                     # pair (1, 0) stored as (image 0, image 1)
    truth = {(int(a), int(b)) for (a, b), ok in zip(matches, is_in) if ok}
    wrong = {(int(a), int(b)) for (a, b), ok in zip(matches, is_in) if not ok}
    stored = {(int(a), int(b)) for a, b in got}
    assert len(stored & truth) >= 0.95 * n_in
    assert len(stored & wrong) <= 0.05 * n_out                     # a random wrong match fits the epipolar geometry by chance ~1 %
    # the explicit opt-out stores everything the descriptor stage produced
    os.remove(db + "-journal") if os.path.exists(db + "-journal") else None
    con = sqlite3.connect(db)
    con.execute("delete from matches")
    con.commit()
    con.close()
    out = subprocess.run([EXE, "match", db, "0", "noverify"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    assert len(_read_matches(db)[1]) >= 0.95 * n


@pytest.mark.gpu
def test_sequential_feature_matcher(tmp_path):
    _need_exe()
    rng = np.random.default_rng(5)
    db = str(tmp_path / "s.db")
    sizes = [400, 500, 450, 300, 350]
    descs, _ = _make_db(db, rng, sizes)
    out = subprocess.run([EXE, "seq", db, "noverify"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = _read_matches(db)
    want = {10000 * j + i for i in range(1, 5) for j in range(max(0, i - 3), i)}          # overlap = 3
    assert set(got) == want
    # integral (un-normalised) descriptors: FilterMatchesByDistance compares the RAW distances with max_distance 0.7, exactly
    # as the reference would on such a database (FeatureMatching.cpp:49) — practically only duplicates survive
    em, _ = mo.match_image_pair(descs[4], descs[2], 0.8, 0.7, True, True)
    np.testing.assert_array_equal(got[20004], em[:, ::-1].reshape(-1, 2))
    em_all, _ = mo.match_image_pair(descs[4], descs[2], 0.8, -1.0, True, True)
    assert len(em) < len(em_all)


@pytest.mark.gpu
def test_feature_utils_two_mats(tmp_path):
    _need_exe()
    rng = np.random.default_rng(9)
    a, b = _sift_like(rng, 333), _sift_like(rng, 500)
    b[rng.permutation(500)[:100]] = a[rng.permutation(333)[:100]]
    (tmp_path / "a.u8").write_bytes(a.tobytes())
    (tmp_path / "b.u8").write_bytes(b.tobytes())
    out = subprocess.run([EXE, "two", str(tmp_path / "a.u8"), "333", str(tmp_path / "b.u8"), "500", str(tmp_path / "o.txt")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = (tmp_path / "o.txt").read_text().split("\n")
    n1 = int(lines[0])
    m1 = np.array([l.split() for l in lines[1:1 + n1]], np.float64).reshape(-1, 3)
    n2 = int(lines[1 + n1])
    m2 = np.array([l.split() for l in lines[2 + n1:2 + n1 + n2]], np.float64).reshape(-1, 3)
    e1, d1 = mo.compute_matches(a, b, 0.8)
    e2, d2 = mo.compute_cross_matches(a, b, 0.8, True)
    np.testing.assert_array_equal(m1[:, :2].astype(np.int32), e1)
    np.testing.assert_allclose(m1[:, 2], d1, rtol=1e-7)
    np.testing.assert_array_equal(m2[:, :2].astype(np.int32), e2)


@pytest.mark.gpu
def test_ceres_bundel_optimizer_dropin(tmp_path):
    _need_exe()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
    name = "ring16"
    cams, pts = g[f"{name}/cams"], g[f"{name}/pts"]
    cx, cy = 1080.0, 720.0
    xy = g[f"{name}/obs_uv"] + [cx, cy]
    oc, op = g[f"{name}/obs_cam"].astype(np.int32), g[f"{name}/obs_pt"].astype(np.int32)
    const = np.nonzero(g[f"{name}/cam_const"])[0].astype(np.int32)
    fin = tmp_path / "in.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("4i", len(cams), len(pts), len(oc), len(const)))
        f.write(struct.pack("4d", float(g[f"{name}/fx"]), float(g[f"{name}/fy"]), cx, cy))
        for arr in (cams, pts, xy):
            f.write(np.ascontiguousarray(arr, np.float64).tobytes())
        for arr in (oc, op, const):
            f.write(arr.tobytes())
    out = subprocess.run([EXE, "ba", str(fin), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "Bundle Adjustment statistics" in out.stdout and "Final RMSE" in out.stdout
    raw = np.fromfile(tmp_path / "out.bin", np.float64)
    ok, before, after, c0, c1, iters = raw[:6]
    costs = g[f"{name}/lm_costs"]
    assert ok == 1.0 and after < before
    assert abs(c0 - costs[0]) <= 1e-9 * costs[0] and abs(c1 - costs[-1]) <= 1e-5 * costs[-1]
    new_cams = raw[6:6 + cams.size].reshape(-1, 6)
    new_pts = raw[6 + cams.size:6 + cams.size + pts.size].reshape(-1, 3)
    # second Optimize with Parameters::refine_focal_length = true from a 2 % wrong K: focal pulled back, error lowered
    before2, after2, fx_ratio, fy_ratio = raw[6 + cams.size + pts.size:6 + cams.size + pts.size + 4]
    # (with one fixed camera the scene scale can trade against the focal length: the ratios are checked loosely)
    assert after2 < 0.2 * before2 and abs(fx_ratio - 1.0) < 1.5e-2 and abs(fy_ratio - 1.0) < 1.5e-2, (before2, after2, fx_ratio, fy_ratio)
    np.testing.assert_array_equal(new_cams[const], cams[const])
    r = bo.residuals_only(new_cams, new_pts, g[f"{name}/obs_uv"], oc, op, float(g[f"{name}/fx"]), float(g[f"{name}/fy"]))
    assert abs(bo.cost_of(r) - c1) <= 1e-9 * c1
    # BundleData::Debug() = mean over landmarks of the mean reprojection error (BundleData.cpp:9-37)
    nrm = np.linalg.norm(r, axis=1)
    per_pt = np.bincount(op, nrm) / np.bincount(op)
    assert abs(per_pt.mean() - after) < 1e-9


@pytest.mark.gpu
def test_optimizer_keeps_the_device_problem_across_calls(tmp_path):
    """MapBuilder owns ONE CeresBundelOptimizer (MapBuilder.cpp:92) and calls Optimize on a map that changes between the calls.
    The device problem persists (msfm_ba_update): same sparsity pattern -> only values are uploaded; a changed map is analysed
    again into the same object.  Every call must give what a fresh optimizer gives on the same input."""
    _need_exe()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
    name = "ring16"
    cams, pts = g[f"{name}/cams"], g[f"{name}/pts"]
    cx, cy = 1080.0, 720.0
    xy = g[f"{name}/obs_uv"] + [cx, cy]
    oc, op = g[f"{name}/obs_cam"].astype(np.int32), g[f"{name}/obs_pt"].astype(np.int32)
    const = np.nonzero(g[f"{name}/cam_const"])[0].astype(np.int32)
    fin = tmp_path / "in.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("4i", len(cams), len(pts), len(oc), len(const)))
        f.write(struct.pack("4d", float(g[f"{name}/fx"]), float(g[f"{name}/fy"]), cx, cy))
        for arr in (cams, pts, xy):
            f.write(np.ascontiguousarray(arr, np.float64).tobytes())
        for arr in (oc, op, const):
            f.write(arr.tobytes())
    out = subprocess.run([EXE, "ba_persist", str(fin), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    raw = np.fromfile(tmp_path / "out.bin", np.float64)
    calls = raw[:21].reshape(3, 7)
    reused, h2d, cost, fresh_cost, fresh_h2d, iters, worst = calls.T
    assert reused.tolist() == [0.0, 1.0, 0.0]
    # the fp32 block sums are accumulated with atomics (order varies from launch to launch): same optimum, not the same bits
    assert (np.abs(cost - fresh_cost) <= 1e-6 * fresh_cost).all(), (cost, fresh_cost)
    assert (worst < 1e-5).all(), worst
    values_only = (cams.size + pts.size + xy.size) * 8
    assert h2d[1] == values_only and h2d[1] < 0.75 * h2d[0] and h2d[0] == fresh_h2d[0]
    assert h2d[2] == fresh_h2d[2] < h2d[0]
    n_keep, n_pts_now, kept_sum, err_sum = raw[21:25]
    assert n_pts_now == len(pts) - len(range(0, len(pts), 7)) and n_keep == (op % 7 != 0).sum()
    assert kept_sum >= 0.95 * n_keep and 0 < err_sum / n_pts_now < 2.0


def _two_view_scene(seed, n_in, n_out, noise_px):
    """3-D points seen by two cameras (NEU intrinsics), `n_out` mismatched pairs appended; float32 pixel coordinates."""
    rng = np.random.default_rng(seed)
    fx = fy = 1449.2752980237
    cx, cy = 1080.0, 720.0
    X = np.c_[rng.uniform(-3, 3, n_in), rng.uniform(-2, 2, n_in), rng.uniform(6, 14, n_in)]
    rvec = np.array([0.03, -0.12, 0.02])
    R = bo.rodrigues_matrix(rvec)
    t = np.array([1.2, 0.1, 0.3])
    def proj(P):
        return np.c_[fx * P[:, 0] / P[:, 2] + cx, fy * P[:, 1] / P[:, 2] + cy]
    p1 = proj(X) + rng.normal(0, noise_px, (n_in, 2))
    p2 = proj(X @ R.T + t) + rng.normal(0, noise_px, (n_in, 2))
    o1 = np.c_[rng.uniform(0, 2160, n_out), rng.uniform(0, 1440, n_out)]
    o2 = np.c_[rng.uniform(0, 2160, n_out), rng.uniform(0, 1440, n_out)]
    truth = np.r_[np.ones(n_in, bool), np.zeros(n_out, bool)]
    perm = rng.permutation(n_in + n_out)
    return np.r_[p1, o1][perm].astype(np.float32), np.r_[p2, o2][perm].astype(np.float32), truth[perm]


@pytest.mark.parametrize("seed,n_in,n_out", [(1, 400, 150), (2, 120, 120), (3, 60, 10)])
def test_filter_matches_geometric_verification(tmp_path, seed, n_in, n_out):
    """FeatureUtils::FilterMatches (F-matrix RANSAC, threshold 3 px, confidence 0.99; FeatureUtils.cpp:176-206) on a
    synthetic two-view scene: keeps the true correspondences, drops the mismatches, and agrees with
    cv2.findFundamentalMat(FM_RANSAC, 3.0, 0.99) — the call the reference makes — on all but a few borderline points."""
    _need_exe()
    p1, p2, truth = _two_view_scene(seed, n_in, n_out, 0.5)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(p1)))
        f.write(p1.tobytes())
        f.write(p2.tobytes())
    out = subprocess.run([EXE, "ransac", str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    mask = np.fromfile(fout, np.uint8).astype(bool)
    assert len(mask) == len(truth)
    recall = (mask & truth).sum() / truth.sum()
    false_pos = (mask & ~truth).sum()
    assert recall >= 0.97, recall
    # a random mismatch survives only if it happens to lie within 3 px of its epipolar lines (~0.5 % here)
    assert false_pos <= max(2, 0.03 * n_out), false_pos
    cv2 = pytest.importorskip("cv2")
    _, cvmask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 3.0, 0.99)
    cvmask = cvmask.ravel().astype(bool)
    assert (cvmask != mask).sum() <= 0.03 * len(mask), int((cvmask != mask).sum())


def test_filter_matches_degenerate_inputs(tmp_path):
    """Fewer than 8 correspondences: nothing survives (OpenCV returns an empty mask below 7; the reference then keeps no
    match)."""
    _need_exe()
    p1, p2, _ = _two_view_scene(5, 6, 0, 0.2)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(p1)))
        f.write(p1.tobytes())
        f.write(p2.tobytes())
    out = subprocess.run([EXE, "ransac", str(fin), str(fout)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    assert not np.fromfile(fout, np.uint8).any()


class _SceneGraphModel:
    """Plain restatement of the reference's SceneGraph (src/Reconstruction/SceneGraph.cpp): per image a list of
    correspondence lists; AddCorrespondences :170-251 (self-matches ignored, invalid / duplicate matches dropped and taken
    back out of the counters), Finalize :88-117, IsTwoViewObservation :285-298."""

    def __init__(self):
        self.corrs, self.ncorr, self.nobs, self.pairs = {}, {}, {}, {}

    def add_image(self, i, n):
        self.corrs[i] = [[] for _ in range(n)]
        self.ncorr[i] = 0
        self.nobs[i] = 0

    def add(self, a, b, matches):
        if a == b:
            return
        pid = 10000 * min(a, b) + max(a, b)                       # Database::ImagePairToPairId (Database.cpp:6)
        self.pairs.setdefault(pid, 0)
        for q, t in matches:
            if not (0 <= q < len(self.corrs[a]) and 0 <= t < len(self.corrs[b])):
                continue
            if (b, t) in self.corrs[a][q]:
                continue
            self.corrs[a][q].append((b, t))
            self.corrs[b][t].append((a, q))
            self.ncorr[a] += 1
            self.ncorr[b] += 1
            self.pairs[pid] += 1

    def finalize(self):
        for i in list(self.corrs):
            self.nobs[i] = sum(1 for c in self.corrs[i] if c)
            if self.nobs[i] == 0:
                del self.corrs[i]

    def dump(self):
        out = [f"images {len(self.corrs)}"]
        for i in sorted(self.corrs):
            out.append(f"image {i} obs {self.nobs[i]} corrs {self.ncorr[i]}")
        for p in sorted(self.pairs):
            out.append(f"pair {p} {self.pairs[p]}")
        return out

    def q(self, i, p):
        c = self.corrs[i][p]
        two = len(c) == 1 and len(self.corrs[c[0][0]][c[0][1]]) == 1
        return f"q {i} {p} has {1 if c else 0} two {1 if two else 0} :" + "".join(f" {a},{b}" for a, b in c)

    def b(self, a, b):
        pid = 10000 * min(a, b) + max(a, b)
        found = [(p, t) for p, cl in enumerate(self.corrs[a]) for (o, t) in cl if o == b]
        return f"b {a} {b} n {self.pairs.get(pid, 0)} :" + "".join(f" {x},{y}" for x, y in found)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scene_graph_against_reference_semantics(tmp_path, seed):
    """SceneGraph (consumer of the matches the M-path writes; CSR storage instead of the reference's nested vectors) driven
    by a random script with duplicates, out-of-range indices, self-matches, repeated pairs, interleaved queries and a
    Finalize: every observable equals the restated reference semantics."""
    _need_exe()
    rng = np.random.default_rng(seed)
    n_img = 7
    npts = {i + 1: int(rng.integers(0, 9)) for i in range(n_img)}
    npts[n_img] = 0                                               # an image without key points
    model = _SceneGraphModel()
    script, expect = [], []
    for i, n in npts.items():
        script.append(f"I {i} {n}")
        model.add_image(i, n)

    def queries():
        for i, n in npts.items():
            if i not in model.corrs:
                continue
            for p in range(n):
                script.append(f"Q {i} {p}")
                expect.append(model.q(i, p))
        for a in model.corrs:
            for b in model.corrs:
                if a != b:
                    script.append(f"B {a} {b}")
                    expect.append(model.b(a, b))
        script.append("D")
        expect.extend(model.dump())

    for rnd in range(3):
        for _ in range(8):
            a, b = (int(x) for x in rng.integers(1, n_img + 1, 2))
            k = int(rng.integers(0, 7))
            m = [(int(rng.integers(-1, npts[a] + 2)), int(rng.integers(-1, npts[b] + 2))) for _ in range(k)]
            if k and rng.random() < 0.5:
                m.append(m[0])                                    # a duplicate inside one call
            script.append(f"M {a} {b} {len(m)} " + " ".join(f"{q} {t}" for q, t in m))
            model.add(a, b, m)
        queries()
    script.append("F")
    model.finalize()
    queries()
    fin, fout = tmp_path / "s.txt", tmp_path / "o.txt"
    fin.write_text("\n".join(script) + "\n")
    out = subprocess.run([EXE, "scenegraph", str(fin), str(fout)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-2000:]
    got = fout.read_text().splitlines()
    assert got == expect


def test_scene_graph_load_from_database(tmp_path):
    """SceneGraph::Load (:11-85) on a database in the reference's schema: every image is a node, pairs below min_num_matches
    are ignored, blobs stored with swapped columns (image_id1 > image_id2, Database.cpp:637-640) come back the right way round."""
    _need_exe()
    db = str(tmp_path / "g.db")
    subprocess.run([EXE, "cpu", db], capture_output=True, text=True, timeout=60, check=True)     # creates 2 images, 3 key points on image 1
    con = sqlite3.connect(db)
    con.execute("DELETE FROM matches")
    con.execute("INSERT INTO keypoints(image_id, rows, cols, data) VALUES (2, 4, 4, ?)", (np.zeros((4, 4), np.float32).tobytes(),))
    m = np.array([[0, 3], [2, 1], [2, 1], [1, 9]], np.int32)     # (point in image 1, point in image 2): one duplicate, one invalid
    con.execute("INSERT INTO matches(pair_id, rows, cols, data) VALUES (?, ?, 2, ?)", (10000 * 1 + 2, len(m), m.tobytes()))
    con.commit()
    con.close()
    fout = tmp_path / "o.txt"
    out = subprocess.run([EXE, "scenegraph_db", db, "3", str(fout)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-2000:]
    assert fout.read_text().splitlines() == ["images 2", "image 1 obs 0 corrs 2", "image 2 obs 0 corrs 2", "pair 10002 2"]
    out = subprocess.run([EXE, "scenegraph_db", db, "5", str(fout)], capture_output=True, text=True, timeout=60)
    assert fout.read_text().splitlines() == ["images 2", "image 1 obs 0 corrs 0", "image 2 obs 0 corrs 0"]
