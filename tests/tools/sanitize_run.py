#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck); checks results too."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np  # noqa: E402

import monocularsfm_b200 as m  # noqa: E402
from oracle import ba_oracle as bo  # noqa: E402
from oracle import match_oracle as mo  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    ctx = m.Context(0)
    sizes = [300, 777, 130]
    imgs = []
    base = rng.integers(0, 256, (800, 128), dtype=np.uint8)
    for k, n in enumerate(sizes):
        x = rng.integers(0, 256, (n, 128), dtype=np.uint8)
        x[:60] = np.clip(base[:60].astype(np.int64) + rng.integers(-2, 3, (60, 128)), 0, 255)
        imgs.append(x)
        ctx.upload(k, x)
    pairs = [(1, 0), (2, 0), (2, 1)]
    off, mt, d = ctx.match_pairs(pairs, m.MatchOptions(0.8, -1.0, True, True))
    for p, (i, j) in enumerate(pairs):
        em, ed = mo.match_image_pair(imgs[i], imgs[j], 0.8, -1.0, True, True)
        assert np.array_equal(mt[off[p]:off[p + 1]], em) and np.array_equal(d[off[p]:off[p + 1]], ed), p
    for mode in (0, 1):
        idx, dist, d2 = ctx.knn2(imgs[0], imgs[1], mode)
        oi, od, _ = mo.knn2(imgs[0], imgs[1])
        assert np.array_equal(idx[:, 0], oi[:, 0]) and np.array_equal(dist, od), mode
    P = bo.make_problem(8, 200, 6, 0)
    ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    s = ba.solve()
    assert s["termination"] == 0, s
    ba.close()
    # long tracks (split tiles), shuffled observations, two constant cameras; with and without the shared focal block
    P = bo.make_long_track_problem(n_cams=80, n_pts=120)
    for focal in (False, True):
        ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"], refine_focal=focal)
        ba.linearize(1e-4)
        s2 = ba.solve()
        assert s2["termination"] == 0, s2
        ba.track_errors()
        ba.close()
    # a narrow-band camera system: the cooperative band Cholesky kernel
    P = bo.make_problem(120, 500, 4, 13)
    ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    s3 = ba.solve()
    assert s3["termination"] == 0 and ba.solver_info()["kind"].startswith("band"), (s3, ba.solver_info())
    # the same object serves the next problems (msfm_ba_update): same pattern (values only), a larger one (the arena grows),
    # a smaller one (the arena is re-used)
    assert ba.update(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"]) is True
    assert ba.solve()["termination"] == 0
    for Q in (bo.make_long_track_problem(n_cams=130, n_pts=700), bo.make_problem(9, 60, 4, 2)):
        assert ba.update(Q["cams"], Q["pts"], Q["obs_uv"], Q["obs_cam"], Q["obs_pt"], Q["cam_const"], Q["fx"], Q["fy"]) is False
        ba.linearize(1e-4)
        assert ba.solve()["termination"] == 0
        ba.filter_stats(4.0)
    ba.close()
    # a WIDE band (ring with tracks spanning up to 64 cameras, the generator of bench.py): 16 tiles below the diagonal, so the
    # back substitution runs on CTA 0 plus 12 helper CTAs (flags, fp64 reductions into zfar)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    import bench
    P = bench.make_ba_problem(400, 3000, 9.2, 7)
    ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    S, rhs, _, _ = ba.linearize(1e-4)
    dc, st = ba.solve_system(1e-4)
    info = ba.solver_info()
    assert st == 0 and info["kind"].startswith("band") and info["band_tiles"] > 4, info
    assert np.abs(S @ dc - rhs).max() <= 1e-6 * np.abs(rhs).max()
    ba.close()
    # batched geometric verification
    k1 = rng.uniform(0, 1000, (300, 2)).astype(np.float32)
    k2 = (k1 + np.float32(5.0)).astype(np.float32)
    ctx.upload_keypoints(900, k1)
    ctx.upload_keypoints(901, k2)
    mm = np.c_[np.arange(300), np.arange(300)].astype(np.int32)
    mask, counts = ctx.verify_pairs([(900, 901)], [0, 300], mm)
    assert counts[0] == mask.sum()
    ctx.close()
    print("sanitize workload ok", len(mt), "matches, BA", s["iterations"], "iterations")


if __name__ == "__main__":
    main()
