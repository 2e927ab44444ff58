"""ctypes binding of include/msfm_b200.h.  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MSFM_OK = 0
MSFM_E_CAPACITY = -4


class MsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"msfm error {code}: {msg}")
        self.code = code


class MatchOptions(C.Structure):
    """msfm_match_options (defaults = the reference's effective defaults, FeatureMatching.h:93-101,
    except max_distance which is off on the u8 scale)."""
    _fields_ = [("max_distance", C.c_double), ("distance_ratio", C.c_float), ("cross_check", C.c_int32),
                ("opencv_quirks", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, distance_ratio=0.8, max_distance=-1.0, cross_check=True, opencv_quirks=True):
        super().__init__(float(max_distance), float(distance_ratio), int(bool(cross_check)),
                         int(bool(opencv_quirks)), 0)


class VerifyOptions(C.Structure):
    """msfm_verify_options; defaults = cv::findFundamentalMat(FM_RANSAC, 3.0, 0.99) as FeatureUtils.cpp:196 calls it."""
    _fields_ = [("threshold", C.c_double), ("confidence", C.c_double), ("max_iters", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, threshold=3.0, confidence=0.99, max_iters=1000):
        super().__init__(float(threshold), float(confidence), int(max_iters), 0)


class BAProblemC(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_pts", C.c_int32), ("n_obs", C.c_int32), ("flags", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cams", C.c_void_p), ("pts", C.c_void_p),
                ("obs_uv", C.c_void_p), ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("cam_const", C.c_void_p)]


class BAOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("verbose", C.c_int32), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double)]


class BASummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("successful_steps", C.c_int32), ("termination", C.c_int32),
                ("num_residuals", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("total_time_s", C.c_double), ("linearize_time_s", C.c_double), ("solve_time_s", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path() -> str:
    # MSFM_B200_LIB: another build of the same library (A/B measurements of kernel variants, tools/dev/build_variants.sh)
    return os.environ.get("MSFM_B200_LIB") or os.path.join(_HERE, "libmsfm_b200.so")


def header_path() -> str:
    return os.path.join(_HERE, "..", "include", "msfm_b200.h")


def exported_symbols():
    """Names of all functions include/msfm_b200.h declares."""
    with open(header_path()) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msfm_[a-z0-9_]+)\s*\(", src)))


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(
            f"{p} is missing: build it with `make` (or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(p)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    P = C.POINTER
    sig = {
        "msfm_init": (C.c_int, [P(vp), C.c_int]),
        "msfm_destroy": (None, [vp]),
        "msfm_last_error": (C.c_char_p, [vp]),
        "msfm_version": (C.c_char_p, []),
        "msfm_sync": (C.c_int, [vp]),
        "msfm_stream": (vp, [vp]),
        "msfm_launch_count": (i64, [vp]),
        "msfm_prof_enable": (C.c_int, [vp, C.c_int]),
        "msfm_prof_reset": (C.c_int, [vp]),
        "msfm_prof_read": (C.c_int, [vp, P(C.c_double), P(i64)]),
        "msfm_desc_upload_u8": (C.c_int, [vp, i32, vp, i32]),
        "msfm_desc_upload_u8_dev": (C.c_int, [vp, i32, vp, i32]),
        "msfm_desc_upload_f32": (C.c_int, [vp, i32, vp, i32, i32]),
        "msfm_desc_upload_raw_f32": (C.c_int, [vp, i32, vp, i32, i32, vp]),
        "msfm_desc_quantised": (C.c_int, [vp, i32]),
        "msfm_desc_count": (C.c_int, [vp, i32]),
        "msfm_desc_release": (C.c_int, [vp, i32]),
        "msfm_desc_release_all": (C.c_int, [vp]),
        "msfm_match_pairs": (C.c_int, [vp, vp, i32, P(MatchOptions), vp, vp, vp, i64, P(i64)]),
        "msfm_match_pairs_dev": (C.c_int, [vp, vp, i32, P(MatchOptions), vp, vp, vp, i64, P(i64)]),
        "msfm_match_knn2_u8": (C.c_int, [vp, vp, i32, vp, i32, i32, vp, vp, vp]),
        "msfm_match_stats": (C.c_int, [vp, P(i64)]),
        "msfm_keypoints_upload": (C.c_int, [vp, i32, vp, i32]),
        "msfm_keypoints_count": (C.c_int, [vp, i32]),
        "msfm_keypoints_release_all": (C.c_int, [vp]),
        "msfm_verify_default_options": (None, [P(VerifyOptions)]),
        "msfm_verify_pairs": (C.c_int, [vp, vp, i32, vp, vp, P(VerifyOptions), vp, vp]),
        "msfm_verify_pairs_dev": (C.c_int, [vp, vp, i32, vp, vp, P(VerifyOptions), vp, vp]),
        "msfm_ba_default_options": (None, [P(BAOptions), i32]),
        "msfm_ba_create": (C.c_int, [vp, P(BAProblemC), P(vp)]),
        "msfm_ba_destroy": (None, [vp]),
        "msfm_ba_update": (C.c_int, [vp, P(BAProblemC), P(i32)]),
        "msfm_ba_last_upload": (C.c_int, [vp, P(i64)]),
        "msfm_ba_sizes": (C.c_int, [vp, P(i64)]),
        "msfm_ba_structure": (C.c_int, [vp, P(i32)]),
        "msfm_ba_get_params": (C.c_int, [vp, vp, vp]),
        "msfm_ba_set_params": (C.c_int, [vp, vp, vp]),
        "msfm_ba_evaluate": (C.c_int, [vp, vp, vp, P(C.c_double)]),
        "msfm_ba_track_errors": (C.c_int, [vp, vp]),
        "msfm_ba_filter_stats": (C.c_int, [vp, C.c_double, vp, vp, vp, vp]),
        "msfm_ba_linearize_focal": (C.c_int, [vp, C.c_double, vp, vp, vp, vp]),
        "msfm_ba_get_focal": (C.c_int, [vp, vp]),
        "msfm_ba_linearize": (C.c_int, [vp, C.c_double, vp, vp, vp, P(C.c_double), P(i32)]),
        "msfm_ba_solve": (C.c_int, [vp, P(BAOptions), P(BASummary)]),
        "msfm_ba_solver_info": (C.c_int, [vp, P(i32)]),
        "msfm_ba_solve_system": (C.c_int, [vp, C.c_double, vp, P(i32)]),
        "msfm_comm_unique_id": (C.c_int, [vp]),
        "msfm_comm_init": (C.c_int, [vp, i32, i32, vp]),
        "msfm_comm_destroy": (C.c_int, [vp]),
        "msfm_comm_allreduce_f64": (C.c_int, [vp, vp, i64, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One msfm_ctx (one per process and device)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.msfm_init(C.byref(h), int(device))
        if rc != MSFM_OK:
            raise MsfmError(rc, (self.lib.msfm_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.msfm_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, allow=()):
        if rc != MSFM_OK and rc not in allow:
            raise MsfmError(rc, (self.lib.msfm_last_error(self.h) or b"").decode())
        return rc

    # ---- context
    def sync(self):
        self._check(self.lib.msfm_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.msfm_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.msfm_launch_count(self.h))

    PROF_NAMES = ["desc_format", "build_units", "match_tile", "resolve", "exact", "compact", "ba_eval", "ba_schur",
                  "ba_other", "ba_comm", "verify", "ba_solve"]

    def prof_enable(self, on=True):
        self._check(self.lib.msfm_prof_enable(self.h, int(on)))

    def prof_reset(self):
        self._check(self.lib.msfm_prof_reset(self.h))

    def prof_read(self):
        n = len(self.PROF_NAMES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._check(self.lib.msfm_prof_read(self.h, ms, cnt))
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(self.PROF_NAMES)}

    # ---- M-path
    def upload(self, image_id: int, desc: np.ndarray):
        desc = np.ascontiguousarray(desc, dtype=np.uint8)
        assert desc.ndim == 2 and desc.shape[1] == 128, desc.shape
        self._check(self.lib.msfm_desc_upload_u8(self.h, image_id, _ptr(desc), desc.shape[0]))

    def upload_f32(self, image_id: int, desc: np.ndarray, always_quantise: bool = False):
        """float32 descriptors as the reference's database stores them; converted to uint8 on the device."""
        desc = np.ascontiguousarray(desc, dtype=np.float32)
        assert desc.ndim == 2 and desc.shape[1] == 128, desc.shape
        self._check(self.lib.msfm_desc_upload_f32(self.h, image_id, _ptr(desc), desc.shape[0], 1 if always_quantise else 0))

    def upload_raw_f32(self, image_id: int, desc: np.ndarray, normalization: str = "l1_root"):
        """Raw SIFT rows: extraction-time normalisation (FeatureExtraction.cpp:143-160) + x512 quantisation on the device.
        Returns the normalised float32 rows the reference would store in its database."""
        desc = np.ascontiguousarray(desc, dtype=np.float32)
        assert desc.ndim == 2 and desc.shape[1] == 128, desc.shape
        out = np.zeros_like(desc)
        kind = {"l1_root": 1, "l2": 2}[normalization]
        self._check(self.lib.msfm_desc_upload_raw_f32(self.h, image_id, _ptr(desc), desc.shape[0], kind, _ptr(out)))
        return out

    def quantised(self, image_id: int) -> bool:
        return bool(self._check(self.lib.msfm_desc_quantised(self.h, image_id), allow=(0, 1)))

    def upload_dev(self, image_id: int, dev_ptr: int, n: int):
        self._check(self.lib.msfm_desc_upload_u8_dev(self.h, image_id, C.c_void_p(dev_ptr), n))

    def desc_count(self, image_id: int) -> int:
        return self._check(self.lib.msfm_desc_count(self.h, image_id), allow=range(0, 1 << 30))

    def release(self, image_id: int):
        self._check(self.lib.msfm_desc_release(self.h, image_id))

    def release_all(self):
        self._check(self.lib.msfm_desc_release_all(self.h))

    def match_pairs(self, pairs, opt: MatchOptions | None = None, capacity: int | None = None, want_dist=True):
        """Returns (offsets int64 [P+1], matches int32 [total,2], dist float32 [total] | None)."""
        opt = opt or MatchOptions()
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        npairs = pairs.shape[0]
        if capacity is None:
            capacity = max(1, sum(self.desc_count(int(p[0])) for p in pairs))
        offsets = np.zeros(npairs + 1, np.int64)
        matches = np.zeros((capacity, 2), np.int32)
        dist = np.zeros(capacity, np.float32) if want_dist else None
        total = C.c_int64(0)
        self._check(self.lib.msfm_match_pairs(self.h, _ptr(pairs), npairs, C.byref(opt), _ptr(offsets), _ptr(matches),
                                              _ptr(dist), capacity, C.byref(total)))
        t = int(total.value)
        return offsets, matches[:t], (dist[:t] if want_dist else None)

    def match_pairs_dev(self, pairs, opt, offsets_ptr: int, matches_ptr: int, dist_ptr: int, capacity: int) -> int:
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        total = C.c_int64(0)
        self._check(self.lib.msfm_match_pairs_dev(self.h, _ptr(pairs), pairs.shape[0], C.byref(opt),
                                                  C.c_void_p(offsets_ptr), C.c_void_p(matches_ptr),
                                                  C.c_void_p(dist_ptr) if dist_ptr else None, capacity,
                                                  C.byref(total)))
        return int(total.value)

    def knn2(self, a: np.ndarray, b: np.ndarray, mode: int = 0):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        na = a.shape[0]
        idx = np.full((na, 2), -1, np.int32)
        dist = np.full((na, 2), np.inf, np.float32)
        d2 = np.full((na, 2), -1, np.int32)
        self._check(self.lib.msfm_match_knn2_u8(self.h, _ptr(a), na, _ptr(b), b.shape[0], mode, _ptr(idx),
                                                _ptr(dist), _ptr(d2)))
        return idx, dist, d2

    # ---- geometric verification
    def upload_keypoints(self, image_id: int, xy: np.ndarray):
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        self._check(self.lib.msfm_keypoints_upload(self.h, image_id, _ptr(xy), xy.shape[0]))

    def verify_pairs(self, pairs, offsets, matches, opt: "VerifyOptions | None" = None):
        """Inlier mask (uint8 per match) and inlier count per pair of the batched F-matrix RANSAC."""
        opt = opt or VerifyOptions()
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        matches = np.ascontiguousarray(matches, dtype=np.int32).reshape(-1, 2)
        mask = np.zeros(max(1, int(offsets[-1])), np.uint8)
        counts = np.zeros(len(pairs), np.int32)
        self._check(self.lib.msfm_verify_pairs(self.h, _ptr(pairs), len(pairs), _ptr(offsets), _ptr(matches), C.byref(opt),
                                               _ptr(mask), _ptr(counts)))
        return mask[:int(offsets[-1])].astype(bool), counts

    # ---- B-path
    def ba_create(self, cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal=False):
        return BAProblem(self, cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal)

    # ---- multi-GPU
    def comm_unique_id(self) -> bytes:
        buf = (C.c_char * 128)()
        self._check(self.lib.msfm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, uid: bytes):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._check(self.lib.msfm_comm_init(self.h, n_ranks, rank, buf))

    def comm_destroy(self):
        self._check(self.lib.msfm_comm_destroy(self.h))

    def comm_allreduce_f64(self, dev_ptr: int, count: int, op: int = 0):
        self._check(self.lib.msfm_comm_allreduce_f64(self.h, C.c_void_p(dev_ptr), count, op))

    def match_stats(self):
        s = (C.c_int64 * 4)()
        self._check(self.lib.msfm_match_stats(self.h, s))
        return {"rows": s[0], "rescans": s[1], "exact_rows": s[2], "units": s[3]}


class BAProblem:
    """msfm_ba: a flattened BundleData resident on the device."""

    def __init__(self, ctx: Context, cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal=False):
        self.ctx = ctx
        self.lib = ctx.lib
        pr, keep = self._describe(cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal)
        h = C.c_void_p()
        ctx._check(self.lib.msfm_ba_create(ctx.h, C.byref(pr), C.byref(h)))
        self.h = h

    def _describe(self, cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal):
        cams = np.ascontiguousarray(cams, np.float64).reshape(-1, 6)
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        obs_uv = np.ascontiguousarray(obs_uv, np.float64).reshape(-1, 2)
        obs_cam = np.ascontiguousarray(obs_cam, np.int32)
        obs_pt = np.ascontiguousarray(obs_pt, np.int32)
        cam_const = np.ascontiguousarray(cam_const, np.uint8)
        self.n_cams, self.n_pts, self.n_obs = len(cams), len(pts), len(obs_cam)
        self.n_free = int((cam_const == 0).sum())
        self.refine_focal = bool(refine_focal)
        pr = BAProblemC(self.n_cams, self.n_pts, self.n_obs, 1 if refine_focal else 0, float(fx), float(fy), _ptr(cams), _ptr(pts),
                        _ptr(obs_uv), _ptr(obs_cam), _ptr(obs_pt), _ptr(cam_const))
        return pr, (cams, pts, obs_uv, obs_cam, obs_pt, cam_const)

    def update(self, cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal=False) -> bool:
        """msfm_ba_update: the next problem on this object.  True if the structure was kept (same sparsity pattern: only
        the values travelled)."""
        pr, keep = self._describe(cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, refine_focal)
        reused = C.c_int32(0)
        self.ctx._check(self.lib.msfm_ba_update(self.h, C.byref(pr), C.byref(reused)))
        return bool(reused.value)

    def last_upload(self):
        info = (C.c_int64 * 2)()
        self.ctx._check(self.lib.msfm_ba_last_upload(self.h, info))
        return {"h2d_bytes": int(info[0]), "reused": bool(info[1])}

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.msfm_ba_destroy(self.h)
        self.h = None

    __del__ = close

    def structure(self):
        info = (C.c_int32 * 8)()
        self.ctx._check(self.lib.msfm_ba_structure(self.h, info))
        return {"n_free": info[0], "n_blocks": info[1], "n_tiles": info[2], "w_max": info[3], "system_bytes": info[4],
                "smem_per_cta": info[5], "tail_f64": info[6], "n_long_tracks": info[7]}

    def get_params(self):
        cams = np.zeros((self.n_cams, 6))
        pts = np.zeros((self.n_pts, 3))
        self.ctx._check(self.lib.msfm_ba_get_params(self.h, _ptr(cams), _ptr(pts)))
        return cams, pts

    def set_params(self, cams=None, pts=None):
        cams = None if cams is None else np.ascontiguousarray(cams, np.float64)
        pts = None if pts is None else np.ascontiguousarray(pts, np.float64)
        self.ctx._check(self.lib.msfm_ba_set_params(self.h, _ptr(cams), _ptr(pts)))

    def evaluate(self, want_r=True, want_J=True):
        r = np.zeros((self.n_obs, 2)) if want_r else None
        J = np.zeros((self.n_obs, 2, 9), np.float32) if want_J else None
        cost = C.c_double(0)
        self.ctx._check(self.lib.msfm_ba_evaluate(self.h, _ptr(r), _ptr(J), C.byref(cost)))
        return r, J, cost.value

    def focal(self):
        f = (C.c_double * 2)()
        self.ctx._check(self.lib.msfm_ba_get_focal(self.h, f))
        return float(f[0]), float(f[1])

    def linearize_focal(self, inv_radius=0.0):
        """Border of the reduced system for a shared focal block: B [6F,2], F [2,2], rhs_f [2], g_f [2]."""
        n6 = 6 * self.n_free
        B = np.zeros((n6, 2))
        F3 = np.zeros(3)
        rf = np.zeros(2)
        gf = np.zeros(2)
        self.ctx._check(self.lib.msfm_ba_linearize_focal(self.h, float(inv_radius), _ptr(B), _ptr(F3), _ptr(rf), _ptr(gf)))
        return B, np.array([[F3[0], F3[1]], [F3[1], F3[2]]]), rf, gf

    def track_errors(self):
        err = np.zeros(self.n_pts)
        self.ctx._check(self.lib.msfm_ba_track_errors(self.h, _ptr(err)))
        return err

    def filter_stats(self, max_reproj_error: float):
        """(obs_keep bool [n_obs], mean error of the kept observations [n_pts], kept count [n_pts], max parallax in degrees [n_pts])."""
        keep = np.zeros(self.n_obs, np.uint8)
        err = np.zeros(self.n_pts)
        kept = np.zeros(self.n_pts, np.int32)
        ang = np.zeros(self.n_pts)
        self.ctx._check(self.lib.msfm_ba_filter_stats(self.h, float(max_reproj_error), _ptr(keep), _ptr(err), _ptr(kept), _ptr(ang)))
        return keep.astype(bool), err, kept, ang

    def linearize(self, inv_radius=0.0, want_S=True):
        n6 = 6 * self.n_free
        S = np.zeros((n6, n6)) if want_S else None
        rhs = np.zeros(n6)
        gc = np.zeros(n6)
        cost = C.c_double(0)
        nf = C.c_int32(0)
        self.ctx._check(self.lib.msfm_ba_linearize(self.h, float(inv_radius), _ptr(S), _ptr(rhs), _ptr(gc),
                                                   C.byref(cost), C.byref(nf)))
        return S, rhs, gc, cost.value

    def default_options(self):
        o = BAOptions()
        self.lib.msfm_ba_default_options(C.byref(o), self.n_cams)
        return o

    def solve_system(self, inv_radius: float):
        """dc [6F] of the damped reduced camera system by the device solver, and its status (0 = positive definite)."""
        dc = np.zeros(6 * self.n_free)
        st = C.c_int32(0)
        self.ctx._check(self.lib.msfm_ba_solve_system(self.h, float(inv_radius), _ptr(dc), C.byref(st)))
        return dc, int(st.value)

    def solver_info(self):
        info = (C.c_int32 * 6)()
        self.ctx._check(self.lib.msfm_ba_solver_info(self.h, info))
        kind = {0: "dense", 1: "block-tridiagonal chain (library)", 2: "band Cholesky (own kernel)"}[info[5]]
        if info[5] == 2:
            return {"kind": kind, "band_cameras": info[0], "tile": info[1], "tile_rows": info[2], "band_tiles": info[3], "n_free": info[4],
                    "n_superblocks": 0}
        return {"kind": kind, "cams_per_superblock": info[0], "superblock_order": info[1], "n_superblocks": info[2], "n_free": info[4]}

    def solve(self, opt: BAOptions | None = None):
        opt = opt or self.default_options()
        s = BASummary()
        self.ctx._check(self.lib.msfm_ba_solve(self.h, C.byref(opt), C.byref(s)))
        return s.as_dict()
