"""B-path oracle (TEST INFRASTRUCTURE — never imported by the product path).

CPU float64 restatement of the bundle-adjustment inner loop the reference delegates to Ceres Solver:

* ``residual_jacobian_jets``  the exact functor of BundleAutoDiffConstantFocalCostFunction
                              (src/Optimizer/CeresBundleOptimizer.cpp:29-53) evaluated with forward-mode dual
                              numbers ("Jets", 9 infinitesimals = rvec|tvec|point), i.e. what
                              ceres::AutoDiffCostFunction<..., 2, 3, 3, 3> (:65) computes.  The rotation is Ceres'
                              published ``AngleAxisRotatePoint`` (include/ceres/rotation.h) including its
                              first-order Taylor branch for theta^2 <= DBL_EPSILON.
* ``build_normal_equations`` / ``schur_reduce``  what Ceres' SchurEliminator produces for the
                              DENSE_SCHUR / SPARSE_SCHUR choices of :264-273 with e-blocks = points (3) and
                              f-blocks = the free cameras (6), constant cameras dropped (:256-260).
* ``lm_solve``                Levenberg-Marquardt with Ceres' documented trust-region rules and the option block of
                              :262-291 (max 100 iterations, function/gradient/parameter tolerances 1e-6/1e-10/1e-8,
                              initial radius 1e4, Jacobi scaling => Marquardt damping diag(J^T J)/radius).

Ceres itself is a third-party dependency that is NOT under /root/reference (README.md:33 pins only "1.10 or higher";
no lockfile, no submodule) and is not installed in this image, and the reference's own tests hold no vectors for
this path: parity with Ceres is therefore UNPINNED.  What pins this restatement instead:
  - tests/test_oracle_ba.py checks the Jet Jacobians against central finite differences and against
    torch.float64 autograd of an independently written projection (cv2.Rodrigues-style rotation matrix), and the
    residual norm against the reference's own second statement of it,
    Projection::CalculateReprojectionError (src/Reconstruction/Projection.cpp:114-133) restated here as
    ``reprojection_error_via_K``;
  - committed golden vectors tests/golden/ba_golden.npz produced by tests/golden/gen_ba_golden.py.
"""
from __future__ import annotations

import numpy as np

DBL_EPS = np.finfo(np.float64).eps


# --------------------------------------------------------------------------------------------- Jets
class Jet:
    """Vectorised dual number: a [N] values, v [N, D] derivatives."""
    __slots__ = ("a", "v")

    def __init__(self, a, v):
        self.a = a
        self.v = v

    @staticmethod
    def const(a, like):
        return Jet(np.broadcast_to(np.asarray(a, np.float64), like.a.shape).copy(), np.zeros_like(like.v))

    def _lift(self, o):
        return o if isinstance(o, Jet) else Jet.const(o, self)

    def __add__(self, o):
        o = self._lift(o)
        return Jet(self.a + o.a, self.v + o.v)

    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o)
        return Jet(self.a - o.a, self.v - o.v)

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        return Jet(self.a * o.a, self.a[:, None] * o.v + o.a[:, None] * self.v)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o)
        inv = 1.0 / o.a
        q = self.a * inv
        return Jet(q, (self.v - q[:, None] * o.v) * inv[:, None])

    def __rtruediv__(self, o):
        return self._lift(o) / self


def jsqrt(x):
    s = np.sqrt(x.a)
    return Jet(s, x.v / (2.0 * s)[:, None])


def jsin(x):
    return Jet(np.sin(x.a), np.cos(x.a)[:, None] * x.v)


def jcos(x):
    return Jet(np.cos(x.a), -np.sin(x.a)[:, None] * x.v)


def _where(mask, a, b):
    return Jet(np.where(mask, a.a, b.a), np.where(mask[:, None], a.v, b.v))


def angle_axis_rotate_point(w, pt):
    """Ceres rotation.h AngleAxisRotatePoint on Jets (lists of 3 Jets each)."""
    theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2]
    big = theta2.a > DBL_EPS
    # main branch (evaluate on a safe copy so the unused lanes do not produce NaNs)
    safe = Jet(np.where(big, theta2.a, 1.0), theta2.v)
    theta = jsqrt(safe)
    c, s = jcos(theta), jsin(theta)
    ti = 1.0 / theta
    wn = [w[0] * ti, w[1] * ti, w[2] * ti]
    wxp = [wn[1] * pt[2] - wn[2] * pt[1], wn[2] * pt[0] - wn[0] * pt[2], wn[0] * pt[1] - wn[1] * pt[0]]
    tmp = (wn[0] * pt[0] + wn[1] * pt[1] + wn[2] * pt[2]) * (1.0 - c)
    main = [pt[k] * c + wxp[k] * s + wn[k] * tmp for k in range(3)]
    # Taylor branch
    wxp2 = [w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]]
    small = [pt[k] + wxp2[k] for k in range(3)]
    return [_where(big, main[k], small[k]) for k in range(3)]


def residual_jacobian_jets(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy):
    """Per observation: r [N,2], J [N,2,9] (columns rvec(3) | tvec(3) | point(3)).
    ``obs_uv`` is already centred by (cx, cy) like CeresBundleOptimizer.cpp:221-222."""
    cams = np.asarray(cams, np.float64)
    pts = np.asarray(pts, np.float64)
    n = len(obs_cam)
    eye = np.eye(9)
    cam_o = cams[obs_cam]
    pt_o = pts[obs_pt]

    def seed(vals, k):
        return Jet(vals.copy(), np.broadcast_to(eye[k], (n, 9)).copy())

    w = [seed(cam_o[:, k], k) for k in range(3)]
    t = [seed(cam_o[:, 3 + k], 3 + k) for k in range(3)]
    X = [seed(pt_o[:, k], 6 + k) for k in range(3)]
    p = angle_axis_rotate_point(w, X)
    p = [p[k] + t[k] for k in range(3)]
    xp = p[0] / p[2]
    yp = p[1] / p[2]
    rx = fx * xp - obs_uv[:, 0]
    ry = fy * yp - obs_uv[:, 1]
    r = np.stack([rx.a, ry.a], 1)
    J = np.stack([rx.v, ry.v], 1)
    return r, J


def residual_jacobian_jets_focal(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy):
    """BundleAutoDiffCostFunction (CeresBundleOptimizer.cpp:76-121; AutoDiffCostFunction<..., 2, 3, 3, 3, 2> at :116): the
    same residual with (fx, fy) as a FOURTH parameter block `focal` shared by every residual (:227-233).
    Returns r [N,2], J [N,2,11] (columns rvec(3) | tvec(3) | point(3) | focal(2))."""
    cams = np.asarray(cams, np.float64)
    pts = np.asarray(pts, np.float64)
    n = len(obs_cam)
    eye = np.eye(11)
    cam_o = cams[obs_cam]
    pt_o = pts[obs_pt]

    def seed(vals, k):
        return Jet(vals.copy(), np.broadcast_to(eye[k], (n, 11)).copy())

    w = [seed(cam_o[:, k], k) for k in range(3)]
    t = [seed(cam_o[:, 3 + k], 3 + k) for k in range(3)]
    X = [seed(pt_o[:, k], 6 + k) for k in range(3)]
    f = [seed(np.full(n, float(fx)), 9), seed(np.full(n, float(fy)), 10)]
    p = angle_axis_rotate_point(w, X)
    p = [p[k] + t[k] for k in range(3)]
    xp = p[0] / p[2]
    yp = p[1] / p[2]
    rx = f[0] * xp - obs_uv[:, 0]
    ry = f[1] * yp - obs_uv[:, 1]
    return np.stack([rx.a, ry.a], 1), np.stack([rx.v, ry.v], 1)


def reduced_system_focal(r, J11, obs_cam, obs_pt, n_cams, n_pts, cam_const, inv_radius):
    """What Ceres' SchurEliminator does with the focal block among the f-blocks: damped normal equations over
    [free cameras (6 each) | focal (2) | points (3 each)], points eliminated.  Returns the reduced matrix
    R [(6F+2),(6F+2)] = [[S, B], [B^T, F]], its right-hand side [6F+2] (R d = rhs), the gradient [6F+2] and the pieces
    back-substitution needs (Hcp [6F+2, P, 3], damped V^-1 [P,3,3], g_p [P,3], free map, F).  Block algebra in numpy."""
    cam_const = np.asarray(cam_const, bool)
    free_idx = np.nonzero(~cam_const)[0]
    fmap = -np.ones(n_cams, np.int64)
    fmap[free_idx] = np.arange(len(free_idx))
    F = len(free_idx)
    nc = 6 * F + 2
    f_o = fmap[obs_cam]
    free = f_o >= 0
    Jc = np.where(free[:, None, None], J11[:, :, :6], 0.0)
    Jp = J11[:, :, 6:9]
    Jf = J11[:, :, 9:11]
    Hcc = np.zeros((nc, nc))
    g = np.zeros(nc)
    # camera diagonal blocks, camera-focal blocks, focal block
    Ucc = np.einsum("nki,nkj->nij", Jc, Jc)
    Ucf = np.einsum("nki,nkj->nij", Jc, Jf)
    for o in np.nonzero(free)[0]:
        k = 6 * f_o[o]
        Hcc[k:k + 6, k:k + 6] += Ucc[o]
        Hcc[k:k + 6, 6 * F:] += Ucf[o]
        Hcc[6 * F:, k:k + 6] += Ucf[o].T
        g[k:k + 6] += Jc[o].T @ r[o]
    Hcc[6 * F:, 6 * F:] = np.einsum("nki,nkj->ij", Jf, Jf)
    g[6 * F:] = np.einsum("nki,nk->i", Jf, r)
    # camera/focal x point blocks and the point blocks
    Hcp = np.zeros((nc, n_pts, 3))
    Wc = np.einsum("nki,nkj->nij", Jc, Jp)
    Wf = np.einsum("nki,nkj->nij", Jf, Jp)
    for o in range(len(obs_cam)):
        if free[o]:
            Hcp[6 * f_o[o]:6 * f_o[o] + 6, obs_pt[o]] += Wc[o]
        Hcp[6 * F:, obs_pt[o]] += Wf[o]
    V = np.zeros((n_pts, 3, 3))
    gp = np.zeros((n_pts, 3))
    np.add.at(V, obs_pt, np.einsum("nki,nkj->nij", Jp, Jp))
    np.add.at(gp, obs_pt, np.einsum("nki,nk->ni", Jp, r))
    d = np.arange(nc)
    Hcc[d, d] += np.maximum(Hcc[d, d], 1e-6) * inv_radius        # Ceres min_lm_diagonal on every parameter
    d3 = np.arange(3)
    V[:, d3, d3] += np.maximum(V[:, d3, d3], 1e-6) * inv_radius
    Vinv = np.linalg.inv(V)
    T = np.einsum("cpi,pij->cpj", Hcp, Vinv)
    R = Hcc - np.einsum("cpj,dpj->cd", T, Hcp)
    rhs = -(g - np.einsum("cpj,pj->c", T, gp))
    return R, rhs, g, (Hcp, Vinv, gp, fmap, F)


def lm_solve_focal(cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, max_iters=100, function_tol=1e-6,
                   gradient_tol=1e-10, parameter_tol=1e-8, initial_radius=1e4):
    """lm_solve with the shared focal block (refine_focal_length = true): identical trust-region rules, dense algebra."""
    cams = np.array(cams, np.float64)
    pts = np.array(pts, np.float64)
    focal = np.array([fx, fy], np.float64)
    n_cams, n_pts = len(cams), len(pts)
    cam_const = np.asarray(cam_const, bool)
    radius, decrease = initial_radius, 2.0
    r, J = residual_jacobian_jets_focal(cams, pts, obs_uv, obs_cam, obs_pt, focal[0], focal[1])
    cost = cost_of(r)
    initial, costs, converged, it = cost, [cost], False, 0
    while it < max_iters:
        it += 1
        R, rhs, gcf, (Hcp, Vinv, gp, fmap, F) = reduced_system_focal(r, J, obs_cam, obs_pt, n_cams, n_pts, cam_const, 1.0 / radius)
        gmax = max(np.abs(gcf).max(), np.abs(gp).max())
        if gmax <= gradient_tol:
            converged = True
            break
        try:
            dcf = np.linalg.solve(R, rhs)
        except np.linalg.LinAlgError:
            radius /= decrease
            decrease *= 2
            continue
        dp = -np.einsum("pij,pj->pi", Vinv, gp + np.einsum("cpi,c->pi", Hcp, dcf))
        dc = np.zeros((n_cams, 6))
        dc[fmap >= 0] = dcf[:6 * F].reshape(-1, 6)
        df = dcf[6 * F:]
        step_norm = np.sqrt((dcf ** 2).sum() + (dp ** 2).sum())
        x_norm = np.sqrt((cams[~cam_const] ** 2).sum() + (focal ** 2).sum() + (pts ** 2).sum())
        if step_norm <= parameter_tol * (x_norm + parameter_tol):
            converged = True
            break
        Jc = J[:, :, :6].copy()
        Jc[cam_const[obs_cam]] = 0
        Jd = (np.einsum("nki,ni->nk", Jc, dc[obs_cam]) + np.einsum("nki,ni->nk", J[:, :, 6:9], dp[obs_pt]) + J[:, :, 9:11] @ df)
        model_decrease = -(float((r * Jd).sum()) + 0.5 * float((Jd * Jd).sum()))
        ncams, npts, nfocal = cams + dc, pts + dp, focal + df
        new_cost = cost_of(residuals_only(ncams, npts, obs_uv, obs_cam, obs_pt, nfocal[0], nfocal[1]))
        rho = (cost - new_cost) / model_decrease if model_decrease > 0 else -1.0
        if rho > 1e-3:
            cams, pts, focal = ncams, npts, nfocal
            dcost = cost - new_cost
            cost = new_cost
            costs.append(cost)
            radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16)
            decrease = 2.0
            r, J = residual_jacobian_jets_focal(cams, pts, obs_uv, obs_cam, obs_pt, focal[0], focal[1])
            if dcost <= function_tol * cost:
                converged = True
                break
        else:
            radius /= decrease
            decrease *= 2
            if radius < 1e-32:
                break
    return {"cams": cams, "pts": pts, "focal": focal, "iterations": it, "initial_cost": initial, "final_cost": cost,
            "converged": converged, "costs": costs}


def residuals_only(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy):
    """Plain float64 evaluation of the same functor (no derivatives)."""
    cams = np.asarray(cams, np.float64)
    w = cams[obs_cam, :3]
    t = cams[obs_cam, 3:]
    X = np.asarray(pts, np.float64)[obs_pt]
    th2 = (w * w).sum(1)
    big = th2 > DBL_EPS
    th = np.sqrt(np.where(big, th2, 1.0))
    c, s = np.cos(th), np.sin(th)
    wn = w / th[:, None]
    wxp = np.cross(wn, X)
    tmp = (wn * X).sum(1) * (1.0 - c)
    main = X * c[:, None] + wxp * s[:, None] + wn * tmp[:, None]
    small = X + np.cross(w, X)
    p = np.where(big[:, None], main, small) + t
    return np.stack([fx * p[:, 0] / p[:, 2] - obs_uv[:, 0], fy * p[:, 1] / p[:, 2] - obs_uv[:, 1]], 1)


def rodrigues_matrix(rvec):
    """cv::Rodrigues(rvec) -> R (used by Map::UpdateFromBAData, Map.cpp:1175-1206, and Projection.cpp)."""
    rvec = np.asarray(rvec, np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-300:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def reprojection_error_via_K(rvec, tvec, X, xy, K):
    """Projection::CalculateReprojectionError (Projection.cpp:114-133): ||dehom(K [R|t] X) - xy||_2 in pixels."""
    R = rodrigues_matrix(rvec)
    P = K @ np.hstack([R, np.asarray(tvec, np.float64).reshape(3, 1)])
    h = P @ np.append(np.asarray(X, np.float64), 1.0)
    return float(np.linalg.norm(h[:2] / h[2] - np.asarray(xy, np.float64)))


def track_errors(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy, cx=0.0, cy=0.0):
    """Map::ComputeTrackError (src/Reconstruction/Map.cpp:1834-1846): per point, the mean over its track of
    Projection::CalculateReprojectionError (Projection.cpp:114-133) = ||dehom(K [R|t] X) - xy||_2, with R = Rodrigues(rvec)
    as Map::UpdateFromBAData builds it (:1186).  obs_uv are centred by (cx, cy) like the optimizer's inputs
    (CeresBundleOptimizer.cpp:221-222); K is rebuilt with that principal point.  Plain loops: small cases only."""
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    n_pts = len(pts)
    s = np.zeros(n_pts)
    cnt = np.zeros(n_pts)
    for o in range(len(obs_cam)):
        c, p = int(obs_cam[o]), int(obs_pt[o])
        xy = np.asarray(obs_uv[o], np.float64) + np.array([cx, cy])
        s[p] += reprojection_error_via_K(cams[c, :3], cams[c, 3:], pts[p], xy, K)
        cnt[p] += 1
    return np.where(cnt > 0, s / np.maximum(cnt, 1), 0.0)


# --------------------------------------------------------------------------------------------- normal equations
def filter_stats(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy, max_reproj_error):
    """Map::FilterPoints3DWithLargeReprojectionError / ...SmallTriangulationAngle restated (src/Reconstruction/Map.cpp:804-917,
    Projection.cpp:6-19, 114-133, 149-194): keep flag per observation, mean error of the kept observations and their number
    per point, largest parallax angle (degrees) over the camera pairs of every point."""
    cams, pts = np.asarray(cams, np.float64), np.asarray(pts, np.float64)
    n_pts = len(pts)
    Rm = np.stack([rodrigues_matrix(c[:3]) for c in cams])
    p_cam = np.einsum("nij,nj->ni", Rm[obs_cam], pts[obs_pt]) + cams[obs_cam, 3:]
    r = residuals_only(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy)
    err = np.sqrt((r * r).sum(1))
    keep = (p_cam[:, 2] > 0) & ~(err > max_reproj_error)
    kept = np.bincount(obs_pt[keep], minlength=n_pts)
    mean = np.bincount(obs_pt[keep], err[keep], n_pts) / np.maximum(kept, 1)
    centers = -np.einsum("nji,nj->ni", Rm, cams[:, 3:])
    ang = np.zeros(n_pts)
    order = np.argsort(obs_pt, kind="stable")
    starts = np.r_[0, np.cumsum(np.bincount(obs_pt, minlength=n_pts))]
    for p in range(n_pts):
        cs = obs_cam[order[starts[p]:starts[p + 1]]]
        for i in range(len(cs)):
            for j in range(i):
                o1, o2 = centers[cs[i]], centers[cs[j]]
                b, r1, r2 = np.linalg.norm(o1 - o2), np.linalg.norm(pts[p] - o1), np.linalg.norm(pts[p] - o2)
                with np.errstate(invalid="ignore"):
                    a = abs(np.arccos((r1 * r1 + r2 * r2 - b * b) / (2 * r1 * r2)))
                a = 0.0 if np.isnan(a) else min(a, np.pi - a) * 180 / np.pi
                ang[p] = max(ang[p], a)
    return keep, mean, kept, ang


def cost_of(r):
    return 0.5 * float((r * r).sum())


def build_normal_equations(r, J, obs_cam, obs_pt, n_cams, n_pts, cam_const):
    """Dense blocks of J^T J and J^T r.  Returns U [Nc,6,6], gc [Nc,6], V [Np,3,3], gp [Np,3],
    W [Nobs,6,3] (per observation camera-point block).  Constant cameras get zero U/gc/W rows."""
    Jc = J[:, :, :6].copy()
    Jp = J[:, :, 6:]
    free = ~np.asarray(cam_const, bool)[obs_cam]
    Jc[~free] = 0.0
    U = np.zeros((n_cams, 6, 6))
    gc = np.zeros((n_cams, 6))
    V = np.zeros((n_pts, 3, 3))
    gp = np.zeros((n_pts, 3))
    np.add.at(U, obs_cam, np.einsum("nki,nkj->nij", Jc, Jc))
    np.add.at(gc, obs_cam, np.einsum("nki,nk->ni", Jc, r))
    np.add.at(V, obs_pt, np.einsum("nki,nkj->nij", Jp, Jp))
    np.add.at(gp, obs_pt, np.einsum("nki,nk->ni", Jp, r))
    W = np.einsum("nki,nkj->nij", Jc, Jp)
    return U, gc, V, gp, W


def schur_reduce(U, gc, V, gp, W, obs_cam, obs_pt, cam_const, inv_radius):
    """Reduced camera system with Marquardt damping D^2 = diag(J^T J) * inv_radius on every parameter
    (what Ceres' LM strategy + Jacobi scaling amounts to, see module docstring).
    Returns S [6F,6F], rhs [6F] over the F free cameras in index order, plus Vinv_damped and the free map.
    Sign convention: S * dc = rhs with rhs = -(gc - W V^-1 gp);  dp = -V^-1 (gp + W^T dc)."""
    n_cams = U.shape[0]
    free_idx = np.nonzero(~np.asarray(cam_const, bool))[0]
    fmap = -np.ones(n_cams, np.int64)
    fmap[free_idx] = np.arange(len(free_idx))
    Vd = V.copy()
    di = np.arange(3)
    Vd[:, di, di] += np.maximum(V[:, di, di], 1e-6) * inv_radius     # Ceres min_lm_diagonal
    Vinv = np.linalg.inv(Vd)
    Ud = U.copy()
    d6 = np.arange(6)
    Ud[:, d6, d6] += np.maximum(U[:, d6, d6], 1e-6) * inv_radius
    F = len(free_idx)
    S = np.zeros((6 * F, 6 * F))
    rhs = np.zeros(6 * F)
    for k, c in enumerate(free_idx):
        S[6 * k:6 * k + 6, 6 * k:6 * k + 6] = Ud[c]
        rhs[6 * k:6 * k + 6] = -gc[c]
    Y = np.einsum("nij,njk->nik", W, Vinv[obs_pt])            # W V^-1  [N,6,3]
    # group observations by point
    order = np.argsort(obs_pt, kind="stable")
    op = obs_pt[order]
    starts = np.nonzero(np.r_[True, op[1:] != op[:-1]])[0]
    ends = np.r_[starts[1:], len(op)]
    for s, e in zip(starts, ends):
        ids = order[s:e]
        ids = ids[fmap[obs_cam[ids]] >= 0]
        if len(ids) == 0:
            continue
        f = fmap[obs_cam[ids]]
        p = obs_pt[ids[0]]
        YW = np.einsum("aij,bkj->abik", Y[ids], W[ids])          # Y_a W_b^T
        Yg = Y[ids] @ gp[p]
        for a in range(len(ids)):
            rhs[6 * f[a]:6 * f[a] + 6] += Yg[a]
            for b in range(len(ids)):
                S[6 * f[a]:6 * f[a] + 6, 6 * f[b]:6 * f[b] + 6] -= YW[a, b]
    return S, rhs, Vinv, fmap


def back_substitute(dc_free, fmap, Vinv, gp, W, obs_cam, obs_pt, n_pts):
    """dp = -V^-1 (gp + sum_obs W^T dc)."""
    n_cams = len(fmap)
    dc = np.zeros((n_cams, 6))
    dc[fmap >= 0] = dc_free.reshape(-1, 6)
    t = gp.copy()
    np.add.at(t, obs_pt, np.einsum("nij,ni->nj", W, dc[obs_cam]))
    return dc, -np.einsum("pij,pj->pi", Vinv, t)


def lm_solve(cams, pts, obs_uv, obs_cam, obs_pt, cam_const, fx, fy, max_iters=100, function_tol=1e-6,
             gradient_tol=1e-10, parameter_tol=1e-8, initial_radius=1e4, verbose=False):
    """Levenberg-Marquardt with Ceres' trust-region rules (LevenbergMarquardtStrategy + TrustRegionMinimizer):
    radius <- radius / max(1/3, 1 - (2 rho - 1)^3) on success (capped at 1e16), radius <- radius / decrease_factor
    with decrease_factor doubling on failure; step accepted when rho > 1e-3; termination by
    |delta cost| <= function_tol * cost (CONVERGENCE), max |g| <= gradient_tol, |step| <= parameter_tol (|x| + parameter_tol).
    Returns dict(cams, pts, iterations, initial_cost, final_cost, converged, costs)."""
    cams = np.array(cams, np.float64)
    pts = np.array(pts, np.float64)
    n_cams, n_pts = len(cams), len(pts)
    cam_const = np.asarray(cam_const, bool)
    radius = initial_radius
    decrease = 2.0
    r, J = residual_jacobian_jets(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy)
    cost = cost_of(r)
    initial = cost
    costs = [cost]
    converged = False
    it = 0
    while it < max_iters:
        it += 1
        U, gc, V, gp, W = build_normal_equations(r, J, obs_cam, obs_pt, n_cams, n_pts, cam_const)
        gmax = max(np.abs(gc[~cam_const]).max() if (~cam_const).any() else 0.0, np.abs(gp).max())
        if gmax <= gradient_tol:
            converged = True
            break
        S, rhs, Vinv, fmap = schur_reduce(U, gc, V, gp, W, obs_cam, obs_pt, cam_const, 1.0 / radius)
        try:
            dc_free = np.linalg.solve(S, rhs)
        except np.linalg.LinAlgError:
            radius /= decrease
            decrease *= 2
            continue
        dc, dp = back_substitute(dc_free, fmap, Vinv, gp, W, obs_cam, obs_pt, n_pts)
        step_norm = np.sqrt((dc ** 2).sum() + (dp ** 2).sum())
        x_norm = np.sqrt((cams[~cam_const] ** 2).sum() + (pts ** 2).sum())
        if step_norm <= parameter_tol * (x_norm + parameter_tol):
            converged = True
            break
        # model decrease  -(g^T d + 1/2 d^T J^T J d)
        Jc = J[:, :, :6].copy()
        Jc[cam_const[obs_cam]] = 0
        Jd = np.einsum("nki,ni->nk", Jc, dc[obs_cam]) + np.einsum("nki,ni->nk", J[:, :, 6:], dp[obs_pt])
        model_decrease = -(float((r * Jd).sum()) + 0.5 * float((Jd * Jd).sum()))
        ncams, npts = cams + dc, pts + dp
        rn = residuals_only(ncams, npts, obs_uv, obs_cam, obs_pt, fx, fy)
        new_cost = cost_of(rn)
        rho = (cost - new_cost) / model_decrease if model_decrease > 0 else -1.0
        if verbose:
            print(f"it {it} cost {cost:.9e} -> {new_cost:.9e} rho {rho:.3f} radius {radius:.3e}")
        if rho > 1e-3:
            cams, pts = ncams, npts
            dcost = cost - new_cost
            cost = new_cost
            costs.append(cost)
            radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16)
            decrease = 2.0
            r, J = residual_jacobian_jets(cams, pts, obs_uv, obs_cam, obs_pt, fx, fy)
            if dcost <= function_tol * cost:
                converged = True
                break
        else:
            radius /= decrease
            decrease *= 2
            if radius < 1e-32:
                break
    return {"cams": cams, "pts": pts, "iterations": it, "initial_cost": initial, "final_cost": cost,
            "converged": converged, "costs": costs}


# --------------------------------------------------------------------------------------------- synthetic problems
def make_problem(n_cams, n_pts, mean_track, seed, noise_px=0.5, perturb=True):
    """Synthetic BA graph of SURVEY.md §8d: cameras on a ring of radius 10 looking inward (small pose jitter), camera 0
    = identity orientation (rvec = 0, like the reference's first camera, Initializer.cpp:319), points in a [-3,3]^3 cube,
    each point observed by its k nearest-angle cameras, observations = projection + N(0, noise_px), NEU intrinsics
    (config/NEU.yaml:28-31).  Returns a dict of float64/int32 arrays; obs sorted by point."""
    rng = np.random.default_rng(seed)
    fx = fy = 1449.2752980237
    ang = np.linspace(0, 2 * np.pi, n_cams, endpoint=False)
    centers = np.stack([10 * np.sin(ang), np.zeros(n_cams), -10 * np.cos(ang)], 1)   # cam 0 at (0,0,-10) looking +z
    cams = np.zeros((n_cams, 6))
    for i in range(n_cams):
        # rotation about y by +ang brings the optical axis (+z) towards the origin
        rvec = np.array([0.0, ang[i], 0.0])
        if i > 0:
            rvec = rvec + rng.normal(0, 0.02, 3)
        R = rodrigues_matrix(rvec)
        cams[i, :3] = rvec
        cams[i, 3:] = -R @ centers[i]
    cams[0, :3] = 0.0
    pts = rng.uniform(-3, 3, (n_pts, 3))
    k = np.clip(rng.geometric(1.0 / mean_track, n_pts), 2, n_cams)
    pang = np.arctan2(pts[:, 0], -pts[:, 2])
    obs_cam, obs_pt = [], []
    for p in range(n_pts):
        d = np.abs(((ang - pang[p] + np.pi) % (2 * np.pi)) - np.pi)
        near = np.argsort(d, kind="stable")[:k[p]]
        obs_cam.append(np.sort(near))
        obs_pt.append(np.full(k[p], p))
    obs_cam = np.concatenate(obs_cam).astype(np.int32)
    obs_pt = np.concatenate(obs_pt).astype(np.int32)
    uv = residuals_only(cams, pts, np.zeros((len(obs_cam), 2)), obs_cam, obs_pt, fx, fy)
    uv = uv + rng.normal(0, noise_px, uv.shape)
    cam_const = np.zeros(n_cams, np.uint8)
    cam_const[0] = 1                                   # Map.cpp:1138 — the first registered image stays fixed
    if perturb:
        pts = pts + rng.normal(0, 0.01, pts.shape)
        cams = cams.copy()
        cams[1:] += rng.normal(0, 0.005, (n_cams - 1, 6))
    return {"cams": cams, "pts": pts, "obs_uv": uv, "obs_cam": obs_cam, "obs_pt": obs_pt, "cam_const": cam_const,
            "fx": fx, "fy": fy}


# --------------------------------------------------------------------------------------------- C restatement
def make_long_track_problem(seed=3, n_cams=90, n_pts=300):
    """Ring problem with shuffled observation order inside every point, two constant cameras, a few very long tracks
    (up to 70 views: split tiles) and one point without observations."""
    P = make_problem(n_cams, n_pts, 6, seed)
    rng = np.random.default_rng(seed)
    oc, op, uv = [], [], []
    for p in range(n_pts):
        sel = np.nonzero(P["obs_pt"] == p)[0]
        if p == 17:
            continue                                        # a point nobody observes
        cams = P["obs_cam"][sel]
        if p % 40 == 5:                                     # long track: 33..70 cameras
            k = int(rng.integers(33, 71))
            cams = (cams[0] + np.arange(k)) % n_cams
        cams = rng.permutation(cams)
        oc.append(cams); op.append(np.full(len(cams), p))
    oc = np.concatenate(oc).astype(np.int32); op = np.concatenate(op).astype(np.int32)
    uv = residuals_only(P["cams"], P["pts"], np.zeros((len(oc), 2)), oc, op, P["fx"], P["fy"]) + rng.normal(0, 0.5, (len(oc), 2))
    P = dict(P, obs_cam=oc, obs_pt=op, obs_uv=uv)
    P["cam_const"] = P["cam_const"].copy()
    P["cam_const"][7] = 1
    return P



def c_oracle():
    """ctypes handle of oracle/libba_oracle.so (built by oracle/Makefile) or None."""
    import ctypes as C
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libba_oracle.so")
    if not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.ba_oracle_evaluate.restype = C.c_double
    lib.ba_oracle_evaluate.argtypes = [C.c_int, dp, dp, dp, ip, ip, C.c_double, C.c_double, dp, dp]
    lib.ba_oracle_linearize.restype = C.c_double
    lib.ba_oracle_linearize.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, ip, ip, C.c_int, C.c_double,
                                        C.c_double, C.c_double, dp, dp]
    lib.ba_oracle_linearize_mt.restype = C.c_double
    lib.ba_oracle_linearize_mt.argtypes = [C.c_int] + lib.ba_oracle_linearize.argtypes
    return lib


def c_linearize(P, inv_radius, lib=None, threads=1):
    """One evaluate + Schur-eliminate pass of the C oracle (threads > 1: the multi-threaded variant, an all-cores bound the reference
    itself does not reach — it never sets Ceres' num_threads).  Returns (S, rhs, cost, seconds)."""
    import ctypes as C
    import time
    lib = lib or c_oracle()
    cams = np.ascontiguousarray(P["cams"], np.float64)
    pts = np.ascontiguousarray(P["pts"], np.float64)
    uv = np.ascontiguousarray(P["obs_uv"], np.float64)
    oc = np.ascontiguousarray(P["obs_cam"], np.int32)
    op = np.ascontiguousarray(P["obs_pt"], np.int32)
    const = np.asarray(P["cam_const"]).astype(bool)
    cam_free = -np.ones(len(cams), np.int32)
    cam_free[~const] = np.arange((~const).sum(), dtype=np.int32)
    nf = int((~const).sum())
    S = np.zeros((6 * nf, 6 * nf))
    rhs = np.zeros(6 * nf)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    t0 = time.perf_counter()
    args = (len(cams), len(pts), len(oc), cams.ctypes.data_as(dp), pts.ctypes.data_as(dp),
            uv.ctypes.data_as(dp), oc.ctypes.data_as(ip), op.ctypes.data_as(ip),
            cam_free.ctypes.data_as(ip), nf, float(P["fx"]), float(P["fy"]), float(inv_radius),
            S.ctypes.data_as(dp), rhs.ctypes.data_as(dp))
    # threads < 0: |threads| threads on one shared S with atomic updates (what large systems use; tests force it)
    cost = lib.ba_oracle_linearize(*args) if threads in (0, 1) else lib.ba_oracle_linearize_mt(int(threads), *args)
    return S, rhs, cost, time.perf_counter() - t0
