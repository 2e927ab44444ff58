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
    ba.close()
    ctx.close()
    print("sanitize workload ok", len(mt), "matches, BA", s["iterations"], "iterations")


if __name__ == "__main__":
    main()
