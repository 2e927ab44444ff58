"""CPU tests of the B-path structure analysis (monocularsfm_b200/csrc/ba_tiles.hpp): device order, tiles, split tiles for
long tracks, block structure of the reduced camera system — and a numpy emulation of what the fused linearisation kernel
does with them (per-tile local accumulators indexed by local camera pairs, flushed through the slot tables), checked
against the oracle's dense Schur elimination.  The CUDA kernel itself is checked in tests/test_ba_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ba_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tiling") / "libtiling.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "tools", "tiling_harness.cpp")], check=True)
    lib = C.CDLL(so)
    ip = C.POINTER(C.c_int32)
    lib.tiling_build.restype = C.c_int
    lib.tiling_build.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip]
    lib.tiling_fetch.restype = None
    lib.tiling_fetch.argtypes = [ip, ip, ip, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), ip, ip, ip, ip, ip, ip, C.POINTER(C.c_uint32)]
    return lib


def run_tiling(lib, P, max_obs=512, max_pts=64, max_items=16):
    oc = np.ascontiguousarray(P["obs_cam"], np.int32)
    op = np.ascontiguousarray(P["obs_pt"], np.int32)
    const = np.asarray(P["cam_const"]).astype(bool)
    cam_free = -np.ones(len(const), np.int32)
    cam_free[~const] = np.arange((~const).sum(), dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    sizes = np.zeros(9, np.int32)
    rc = lib.tiling_build(len(const), len(P["pts"]), len(oc), oc.ctypes.data_as(ip), op.ctypes.data_as(ip),
                          cam_free.ctypes.data_as(ip), int((~const).sum()), max_obs, max_pts, max_items, sizes.ctypes.data_as(ip))
    if rc:
        return None
    n_tiles, n_tc, n_marks, n_blocks, w_max, n_items, first_long, n_long, n_runs = [int(x) for x in sizes]
    T = {"pt_order": np.zeros(len(P["pts"]), np.int32), "pt_start": np.zeros(len(P["pts"]) + 1, np.int32),
         "obs_perm": np.zeros(len(oc), np.int32), "obs_lcam": np.zeros(len(oc), np.uint8), "obs_lpt": np.zeros(len(oc), np.uint8),
         "first_long": first_long, "n_long": n_long,
         "tiles": np.zeros((n_tiles, 12), np.int32), "runs": np.zeros(n_runs, np.uint32), "items_raw": np.zeros((n_items, 12), np.int32), "tile_cams": np.zeros(n_tc, np.int32),
         "tile_slots": np.zeros(n_marks, np.int32), "blk_row": np.zeros(n_blocks, np.int32),
         "blk_col": np.zeros(n_blocks, np.int32), "w_max": w_max, "cam_free": cam_free}
    lib.tiling_fetch(T["pt_order"].ctypes.data_as(ip), T["pt_start"].ctypes.data_as(ip), T["obs_perm"].ctypes.data_as(ip),
                     T["obs_lcam"].ctypes.data_as(C.POINTER(C.c_uint8)), T["obs_lpt"].ctypes.data_as(C.POINTER(C.c_uint8)),
                     T["tiles"].ctypes.data_as(ip),
                     T["items_raw"].ctypes.data_as(ip), T["tile_cams"].ctypes.data_as(ip), T["tile_slots"].ctypes.data_as(ip), T["blk_row"].ctypes.data_as(ip),
                     T["blk_col"].ctypes.data_as(ip), T["runs"].ctypes.data_as(C.POINTER(C.c_uint32)))
    # struct Item { int32 d; uint16 a0, a1, b0, b1, pad[2]; uint8 lc[32]; }
    raw = T["items_raw"].view(np.uint8).reshape(n_items, 48)
    h = raw[:, 4:16].copy().view(np.uint16).reshape(n_items, 6)
    T["items"] = [{"d": int(T["items_raw"][i, 0]), "a0": int(h[i, 0]), "a1": int(h[i, 1]), "b0": int(h[i, 2]), "b1": int(h[i, 3]),
                   "primary": int(h[i, 0] == 0 and h[i, 2] == 0 and h[i, 3] == 0), "lc": raw[i, 16:48].astype(int)} for i in range(n_items)]
    return T


def tile_units(T, t):
    """(device point, selected observation positions or None, local cameras or None, nA, nB, primary) of every unit of a tile."""
    b, e, ob, no, cb, w, sb, flags = [int(x) for x in t[:8]]
    out = []
    if flags & 1:
        for it in T["items"][b:e]:
            nA, nB = it["a1"] - it["a0"], it["b1"] - it["b0"]
            sel = list(range(it["a0"], it["a1"])) + list(range(it["b0"], it["b1"]))
            out.append((it["d"], sel, it["lc"][:nA + nB].tolist(), nA, nB, bool(it["primary"])))
    else:
        out = [(d, None, None, 0, 0, True) for d in range(b, e)]
    return out


long_track_problem = bo.make_long_track_problem


def tri(la, lb):
    return lb * (lb + 1) // 2 + la


@pytest.mark.parametrize("which", ["ring", "long", "long-tiny-chunks"])
def test_tiling_invariants(harness, which, monkeypatch):
    P = bo.make_problem(40, 500, 8, 1) if which == "ring" else long_track_problem()
    if which == "long-tiny-chunks":
        monkeypatch.setenv("TILING_TEST_CHUNK_PTS", "37")
        monkeypatch.setenv("TILING_TEST_CHUNK_LONG", "2")
    T = run_tiling(harness, P)
    n_pts, n_obs = len(P["pts"]), len(P["obs_cam"])
    assert sorted(T["pt_order"].tolist()) == list(range(n_pts))
    assert sorted(T["obs_perm"].tolist()) == list(range(n_obs))
    oc, op = P["obs_cam"][T["obs_perm"]], P["obs_pt"][T["obs_perm"]]
    for d in range(n_pts):
        s, e = T["pt_start"][d], T["pt_start"][d + 1]
        assert (op[s:e] == T["pt_order"][d]).all()
        assert (np.diff(oc[s:e]) > 0).all()                  # sorted by camera, no duplicates
    assert T["w_max"] <= 32
    covered = np.zeros(n_pts, int)
    pair_count = {}
    for t in T["tiles"]:
        b0_, e0_, ob_, no_, cb, w, sb, flags = [int(x) for x in t[:8]]
        cams = T["tile_cams"][cb:cb + w]
        assert w <= 32 and (np.diff(cams) > 0).all()
        if flags & 1:
            assert e0_ - b0_ <= 16
        rb, nr = int(t[8]), int(t[9])
        runs = T["runs"][rb:rb + nr]
        # work items = (run, round of 32 pairs); the runs (round 0 entries) partition the units in order, and inside a run every
        # unit has the same camera list
        items = runs
        runs = items[(items >> 24) == 0]
        cnt = (runs >> 16) & 0xFF
        kk_ = np.diff(T["pt_start"])
        covered_units = np.zeros(e0_ - b0_, int)
        for rn in runs:
            covered_units[int(rn & 0xFFFF):int(rn & 0xFFFF) + int((rn >> 16) & 0xFF)] += 1
        assert (np.diff(runs & 0xFFFF) > 0).all() and covered_units.max() <= 1
        if not (flags & 1):
            # a unit without a work item has no camera pair (a single observation)
            assert (kk_[b0_:e0_][covered_units == 0] < 2).all() and (kk_[b0_:e0_][covered_units == 1] >= 2).all()
            want_items = sum((int(kk_[b0_ + int(rn & 0xFFFF)]) * (int(kk_[b0_ + int(rn & 0xFFFF)]) - 1) // 2 + 31) // 32 for rn in runs)
            assert len(items) == want_items
            for rn in runs:
                u0, n = int(rn & 0xFFFF), int((rn >> 16) & 0xFF)
                ref = oc[T["pt_start"][b0_ + u0]:T["pt_start"][b0_ + u0 + 1]]
                for u in range(u0 + 1, u0 + n):
                    assert np.array_equal(oc[T["pt_start"][b0_ + u]:T["pt_start"][b0_ + u + 1]], ref)
        if flags & 1:
            pass
        else:
            assert e0_ - b0_ <= 256 and no_ <= 512 and ob_ == T["pt_start"][b0_] and ob_ + no_ == T["pt_start"][e0_]
            assert (T["obs_lpt"][ob_:ob_ + no_] == np.repeat(np.arange(e0_ - b0_), np.diff(T["pt_start"][b0_:e0_ + 1]))).all()
        for d, sel, lc, nA, nB, primary in tile_units(T, t):
            s0, e0 = int(T["pt_start"][d]), int(T["pt_start"][d + 1])
            if sel is None:
                assert e0 - s0 <= 32
                covered[d] += 1
                assert (cams[T["obs_lcam"][s0:e0]] == oc[s0:e0]).all()
            else:                                            # an item: groups of <= 16 observations of a long track
                assert e0 - s0 > 32 and nA <= 16 and nB <= 16
                assert (cams[lc] == oc[[s0 + x for x in sel]]).all() and (np.diff(lc) > 0).all()
                covered[d] += int(primary)
                A, B = sel[:nA], sel[nA:]
                prs = [(x, y) for x in A for y in B] if B else [(x, y) for x in A for y in A if x <= y]
                for x, y in prs:
                    pair_count[(d, x, y)] = pair_count.get((d, x, y), 0) + 1
    k = np.diff(T["pt_start"])
    assert (covered[k > 0] == 1).all() and (covered[k == 0] == 0).all()
    for d in np.nonzero(k > 32)[0]:                           # every pair (and diagonal) of a long track exactly once
        kk = int(k[d])
        assert all(pair_count.get((int(d), x, y), 0) == 1 for x in range(kk) for y in range(x, kk))
    # block structure == co-observing free camera pairs (+ the full diagonal)
    cf = T["cam_free"]
    want = {(f, f) for f in range(int((cf >= 0).sum()))}
    for d in range(n_pts):
        f = cf[oc[T["pt_start"][d]:T["pt_start"][d + 1]]]
        f = f[f >= 0]
        want |= {(int(a), int(b)) for a in f for b in f if a <= b}
    got = list(zip(T["blk_row"].tolist(), T["blk_col"].tolist()))
    assert got == sorted(want)


def test_items_of_neighbouring_long_tracks_are_packed(harness):
    """Ring problem with many long tracks (the BASELINE configs[3] generator, smaller): item tiles hold many items each."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    P = bench.make_ba_problem(128, 20000, 10.0, 4321)
    T = run_tiling(harness, P, 512, 256)
    split = (T["tiles"][:, 7] & 1) == 1
    assert len(T["items"]) > 1000
    # a cross item holds 32 cameras = the whole tile budget, so only long tracks that start at the same camera share a tile
    assert split.sum() * 3 <= len(T["items"])
    nobs = T["tiles"][~split, 3]
    assert nobs.mean() > 250 and np.median(nobs) > 400   # most normal tiles fill their 512 observation slots (not at the ring's wrap-around)
    assert T["n_long"] == int((np.diff(T["pt_start"]) > 32).sum())
    # identical camera lists are adjacent: the pair-weighted mean run length is well above 1 on this ring problem
    k = np.diff(T["pt_start"])
    w = tot = 0
    for t in T["tiles"][~split]:
        for rn in T["runs"][int(t[8]):int(t[8]) + int(t[9])]:
            if rn >> 24:
                continue
            kk, n = int(k[int(t[0]) + int(rn & 0xFFFF)]), int((rn >> 16) & 0xFF)
            w += kk * (kk - 1) // 2 * n * n; tot += kk * (kk - 1) // 2 * n
    assert w / tot > 1.5, w / tot


def test_duplicate_camera_rejected(harness):
    P = bo.make_problem(6, 20, 3, 0)
    oc = P["obs_cam"].copy()
    oc[1] = oc[0]
    assert run_tiling(harness, dict(P, obs_cam=oc)) is None


@pytest.mark.parametrize("which", ["ring", "long", "long-tiny-chunks"])
def test_tile_accumulation_emulation_matches_dense_schur(harness, which, monkeypatch):
    """What fused_linearize_kernel computes, restated in numpy from the tiling tables: per tile a local block triangle over
    the local cameras, block (x, y) -= Jc_x^T (Jp_x V^-1 Jp_y^T) Jc_y, diagonal += Jc^T (I - Jp V^-1 Jp^T) Jc, flushed to
    the global block list through tile_slots.  Must equal the oracle's S, rhs."""
    P = bo.make_problem(40, 500, 8, 1) if which == "ring" else long_track_problem()
    if which == "long-tiny-chunks":
        # the tiling runs in independent chunks of the device order that are concatenated afterwards (16 384 points / 1 024 long
        # tracks in production): chunks of 37 points / 2 long tracks put many chunk boundaries into this small problem
        monkeypatch.setenv("TILING_TEST_CHUNK_PTS", "37")
        monkeypatch.setenv("TILING_TEST_CHUNK_LONG", "2")
    T = run_tiling(harness, P)
    if which == "long-tiny-chunks":
        assert len(T["tiles"]) >= 8 + 4
    inv_radius = 1e-4
    r, J = bo.residual_jacobian_jets(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    U, gc, V, gp, W = bo.build_normal_equations(r, J, P["obs_cam"], P["obs_pt"], len(P["cams"]), len(P["pts"]), P["cam_const"])
    So, rhso, _, _ = bo.schur_reduce(U, gc, V, gp, W, P["obs_cam"], P["obs_pt"], P["cam_const"], inv_radius)
    cf = T["cam_free"]
    nf = int((cf >= 0).sum())
    perm = T["obs_perm"]
    Jc, Jp, rr, oc = J[perm][:, :, :6], J[perm][:, :, 6:], r[perm], P["obs_cam"][perm]
    sblk = np.zeros((len(T["blk_row"]), 6, 6))
    rhs = np.zeros((nf, 6)); udiag = np.zeros((nf, 6))
    for t in T["tiles"]:
        cb, w, sb = int(t[4]), int(t[5]), int(t[6])
        lfree = cf[T["tile_cams"][cb:cb + w]]
        acc = np.zeros((w * (w + 1) // 2, 6, 6))
        camacc = np.zeros((w, 12))
        for d, sel_pos, lc_item, nA, nB, primary in tile_units(T, t):
            s, e = int(T["pt_start"][d]), int(T["pt_start"][d + 1])
            Vp = (Jp[s:e].transpose(0, 2, 1) @ Jp[s:e]).sum(0)
            g = (Jp[s:e].transpose(0, 2, 1) @ rr[s:e, :, None]).sum(0)[:, 0]
            Vp[np.arange(3), np.arange(3)] += np.maximum(np.diag(Vp), 1e-6) * inv_radius
            Vi = np.linalg.inv(Vp)
            if sel_pos is not None:
                sel = [s + x for x in sel_pos]; lc = lc_item
            else:
                sel = list(range(s, e)); lc = T["obs_lcam"][s:e].tolist(); nA, nB = e - s, 0
            Q = [Jp[o] @ Vi for o in sel]
            if nB == 0:
                for x, o in enumerate(sel):
                    if cf[oc[o]] < 0:
                        continue
                    N = np.eye(2) - Q[x] @ Jp[o].T
                    acc[tri(lc[x], lc[x])] += Jc[o].T @ N @ Jc[o]
                    camacc[lc[x], :6] += Jc[o].T @ (Q[x] @ g - rr[o])
                    camacc[lc[x], 6:] += (Jc[o] ** 2).sum(0)
                prs = [(x, y) for y in range(nA) for x in range(y)]
            else:
                prs = [(x, nA + y) for x in range(nA) for y in range(nB)]
            for x, y in prs:
                if cf[oc[sel[x]]] < 0 or cf[oc[sel[y]]] < 0:
                    continue
                assert lc[x] < lc[y]
                acc[tri(lc[x], lc[y])] -= Jc[sel[x]].T @ (Q[x] @ Jp[sel[y]].T) @ Jc[sel[y]]
        slots = T["tile_slots"][sb:sb + w * (w + 1) // 2]
        for b in range(len(slots)):
            if slots[b] >= 0:
                sblk[slots[b]] += acc[b]
            else:
                assert not acc[b].any()
        for l in range(w):
            if lfree[l] >= 0:
                rhs[lfree[l]] += camacc[l, :6]; udiag[lfree[l]] += camacc[l, 6:]
    S = np.zeros((6 * nf, 6 * nf))
    for k, (fa, fb) in enumerate(zip(T["blk_row"], T["blk_col"])):
        S[6 * fa:6 * fa + 6, 6 * fb:6 * fb + 6] = sblk[k]
        S[6 * fb:6 * fb + 6, 6 * fa:6 * fa + 6] = sblk[k].T
    S[np.arange(6 * nf), np.arange(6 * nf)] += np.maximum(udiag.ravel(), 1e-6) * inv_radius
    assert np.abs(S - So).max() <= 1e-10 * np.abs(So).max()
    assert np.abs(rhs.ravel() - rhso).max() <= 1e-9 * np.abs(rhso).max()


def test_analysis_does_not_depend_on_the_thread_count(harness, monkeypatch):
    """The structure analysis runs on the host threads (parallel merge sort of the point keys, tiling in chunks of a constant
    size); one thread and many must give the same tables."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    P = bench.make_ba_problem(200, 60000, 9.2, 99)          # above the thresholds of the parallel paths, several chunks
    monkeypatch.setenv("MSFM_HOST_THREADS", "1")
    A = run_tiling(harness, P, 512, 256)
    monkeypatch.delenv("MSFM_HOST_THREADS")
    B = run_tiling(harness, P, 512, 256)
    for k in ("pt_order", "pt_start", "obs_perm", "obs_lcam", "obs_lpt", "tiles", "runs", "tile_cams", "tile_slots", "blk_row", "blk_col", "items_raw"):
        assert np.array_equal(A[k], B[k]), k
    assert len(A["tiles"]) > 200 and A["n_long"] > 0
