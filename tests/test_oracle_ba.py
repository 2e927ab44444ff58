"""CPU: pin the B-path oracle (float64 Jets restating the reference's Ceres cost functor)."""
import numpy as np
import pytest

from oracle import ba_oracle as bo

NAMES = ["small", "special", "ring16"]


def _prob(g, name):
    return {k: g[f"{name}/{k}"] for k in ("cams", "pts", "obs_uv", "obs_cam", "obs_pt", "cam_const")} | {
        "fx": float(g[f"{name}/fx"]), "fy": float(g[f"{name}/fy"])}


@pytest.fixture(scope="module")
def golden_ba():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ba_golden.npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(golden_ba, name):
    g = golden_ba
    P = _prob(g, name)
    r, J = bo.residual_jacobian_jets(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    np.testing.assert_allclose(r, g[f"{name}/r"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(J, g[f"{name}/J"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", NAMES)
def test_jets_vs_finite_differences(golden_ba, name):
    g = golden_ba
    P = _prob(g, name)
    args = (P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    J = g[f"{name}/J"]
    h = 1e-6
    Jfd = np.zeros_like(J)
    for k in range(6):
        c1, c2 = P["cams"].copy(), P["cams"].copy()
        c1[:, k] += h
        c2[:, k] -= h
        Jfd[:, :, k] = (bo.residuals_only(c1, P["pts"], *args) - bo.residuals_only(c2, P["pts"], *args)) / (2 * h)
    for k in range(3):
        p1, p2 = P["pts"].copy(), P["pts"].copy()
        p1[:, k] += h
        p2[:, k] -= h
        Jfd[:, :, 6 + k] = (bo.residuals_only(P["cams"], p1, *args) - bo.residuals_only(P["cams"], p2, *args)) / (2 * h)
    assert np.abs(J - Jfd).max() <= 2e-6 * np.abs(J).max()


def test_jets_vs_torch_autograd_of_independent_projection(golden_ba):
    """Independent statement: rotation MATRIX from Rodrigues' formula (cv::Rodrigues style), K [R|t] X projection,
    differentiated by torch.float64 autograd.  Cameras in Ceres' Taylor branch are excluded (the matrix form is the
    exact exponential there, the reference's functor is its first-order truncation)."""
    torch = pytest.importorskip("torch")
    g = golden_ba
    P = _prob(g, "special")
    cams = torch.tensor(P["cams"], dtype=torch.float64, requires_grad=True)
    pts = torch.tensor(P["pts"], dtype=torch.float64, requires_grad=True)
    oc = torch.tensor(P["obs_cam"].astype(np.int64))
    op = torch.tensor(P["obs_pt"].astype(np.int64))

    def project(cams, pts):
        w = cams[oc, :3]
        t = cams[oc, 3:]
        X = pts[op]
        th = w.norm(dim=1, keepdim=True).clamp_min(1e-300)
        k = w / th
        zero = torch.zeros_like(k[:, 0])
        K = torch.stack([zero, -k[:, 2], k[:, 1], k[:, 2], zero, -k[:, 0], -k[:, 1], k[:, 0], zero], 1).reshape(-1, 3, 3)
        R = torch.eye(3, dtype=torch.float64) + torch.sin(th)[:, :, None] * K + (1 - torch.cos(th))[:, :, None] * (K @ K)
        p = (R @ X[:, :, None])[:, :, 0] + t
        return torch.stack([P["fx"] * p[:, 0] / p[:, 2], P["fy"] * p[:, 1] / p[:, 2]], 1) - torch.tensor(P["obs_uv"])

    r = project(cams, pts)
    big = (P["cams"][P["obs_cam"], :3] ** 2).sum(1) > 1e-8
    np.testing.assert_allclose(r.detach().numpy()[big], g["special/r"][big], atol=1e-9)
    J = g["special/J"]
    for comp in range(2):
        gc, gp = torch.autograd.grad(r[:, comp].sum(), (cams, pts), retain_graph=True)
        # every camera/point appears in many observations; compare the per-parameter sums
        Jc_sum = np.zeros_like(P["cams"])
        np.add.at(Jc_sum, P["obs_cam"][big], J[big, comp, :6])
        Jp_sum = np.zeros_like(P["pts"])
        np.add.at(Jp_sum, P["obs_pt"], J[:, comp, 6:])
        cam_ok = np.unique(P["obs_cam"][big])
        small_cams = np.unique(P["obs_cam"][~big])
        cam_ok = np.setdiff1d(cam_ok, small_cams)
        np.testing.assert_allclose(gc.numpy()[cam_ok], Jc_sum[cam_ok], rtol=1e-8, atol=1e-7)
        pts_only_big = np.setdiff1d(np.arange(len(P["pts"])), np.unique(P["obs_pt"][~big]))
        np.testing.assert_allclose(gp.numpy()[pts_only_big], Jp_sum[pts_only_big], rtol=1e-8, atol=1e-7)


def test_residual_norm_vs_reference_projection_cpp(golden_ba):
    """Projection::CalculateReprojectionError (Projection.cpp:114-133) is the reference's own second statement of
    ||r||: K [R|t] X, dehomogenise, L2 pixel error."""
    g = golden_ba
    P = _prob(g, "small")
    cx, cy = 1080.0, 720.0
    K = np.array([[P["fx"], 0, cx], [0, P["fy"], cy], [0, 0, 1.0]])
    r = g["small/r"]
    for i in range(0, len(r), 37):
        c, p = P["obs_cam"][i], P["obs_pt"][i]
        e = bo.reprojection_error_via_K(P["cams"][c, :3], P["cams"][c, 3:], P["pts"][p], P["obs_uv"][i] + [cx, cy], K)
        assert abs(e - np.linalg.norm(r[i])) < 1e-9


@pytest.mark.parametrize("name", NAMES)
def test_schur_equals_full_normal_equations(golden_ba, name):
    """Eliminating the points must give the same camera step as solving the full damped system."""
    g = golden_ba
    P = _prob(g, name)
    r, J = g[f"{name}/r"], g[f"{name}/J"]
    nc, npnt = len(P["cams"]), len(P["pts"])
    free = np.nonzero(P["cam_const"] == 0)[0]
    U, gc, V, gp, W = bo.build_normal_equations(r, J, P["obs_cam"], P["obs_pt"], nc, npnt, P["cam_const"])
    lam = 1e-4
    S, rhs, Vinv, fmap = bo.schur_reduce(U, gc, V, gp, W, P["obs_cam"], P["obs_pt"], P["cam_const"], lam)
    np.testing.assert_allclose(S, g[f"{name}/S"], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(S, S.T, rtol=1e-12, atol=1e-8)
    dc = np.linalg.solve(S, rhs)
    # full system
    n = 6 * len(free) + 3 * npnt
    Jfull = np.zeros((2 * len(r), n))
    for i in range(len(r)):
        f = fmap[P["obs_cam"][i]]
        if f >= 0:
            Jfull[2 * i:2 * i + 2, 6 * f:6 * f + 6] = J[i, :, :6]
        p = P["obs_pt"][i]
        Jfull[2 * i:2 * i + 2, 6 * len(free) + 3 * p:6 * len(free) + 3 * p + 3] = J[i, :, 6:]
    H = Jfull.T @ Jfull
    H = H + np.diag(np.maximum(np.diag(H), 1e-6) * lam)
    d = np.linalg.solve(H, -Jfull.T @ r.reshape(-1))
    np.testing.assert_allclose(dc, d[:6 * len(free)], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", NAMES)
def test_lm_converges_and_matches_golden(golden_ba, name):
    g = golden_ba
    P = _prob(g, name)
    res = bo.lm_solve(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    assert res["converged"] and bool(g[f"{name}/lm_converged"])
    np.testing.assert_allclose(res["costs"], g[f"{name}/lm_costs"], rtol=1e-9)
    assert res["final_cost"] < 0.7 * res["initial_cost"]
    # constant cameras did not move
    const = P["cam_const"] == 1
    np.testing.assert_array_equal(res["cams"][const], P["cams"][const])


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_python(golden_ba, name):
    lib = bo.c_oracle()
    if lib is None:
        pytest.skip("oracle/libba_oracle.so not built (make -C oracle)")
    g = golden_ba
    P = _prob(g, name)
    S, rhs, cost, _ = bo.c_linearize(P, 1e-4, lib)
    np.testing.assert_allclose(S, g[f"{name}/S"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(rhs, g[f"{name}/rhs"], rtol=1e-9, atol=1e-7)
    assert abs(cost - float(g[f"{name}/cost"])) < 1e-9 * cost
    # the multi-threaded variant (bench.py's all-cores figure): same sums in another order; both accumulation modes
    for threads in (3, 8, -4):
        S2, rhs2, cost2, _ = bo.c_linearize(P, 1e-4, lib, threads=threads)
        np.testing.assert_allclose(S2, S, rtol=1e-11, atol=1e-9 * np.abs(S).max())
        np.testing.assert_allclose(rhs2, rhs, rtol=1e-11, atol=1e-9 * np.abs(rhs).max())
        assert abs(cost2 - cost) < 1e-12 * cost


@pytest.mark.parametrize("name", NAMES)
def test_track_errors_match_residual_norms(golden_ba, name):
    """oracle.track_errors (Map::ComputeTrackError through K [R|t], Projection.cpp:114-133) == per-point mean of the
    golden residual norms of the optimizer's functor: the reference's two statements of the reprojection error agree."""
    g = golden_ba
    cams, pts = g[f"{name}/cams"], g[f"{name}/pts"]
    obs_cam, obs_pt = g[f"{name}/obs_cam"], g[f"{name}/obs_pt"]
    err = bo.track_errors(cams, pts, g[f"{name}/obs_uv"], obs_cam, obs_pt, float(g[f"{name}/fx"]), float(g[f"{name}/fy"]), 1080.0, 720.0)
    rn = np.linalg.norm(g[f"{name}/r"], axis=1)
    mean = np.bincount(obs_pt, rn, len(pts)) / np.maximum(np.bincount(obs_pt, minlength=len(pts)), 1)
    np.testing.assert_allclose(err, mean, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", NAMES)
def test_focal_jets_vs_finite_differences_and_base_functor(golden_ba, name):
    """BundleAutoDiffCostFunction (:76-121): same residual and same first nine Jacobian columns as the constant-focal functor;
    the two focal columns against central differences (the residual is linear in fx, fy: exact up to rounding)."""
    P = _prob(golden_ba, name)
    r, J = bo.residual_jacobian_jets_focal(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    r0, J0 = bo.residual_jacobian_jets(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    np.testing.assert_array_equal(r, r0)
    np.testing.assert_array_equal(J[:, :, :9], J0)
    h = 1e-3
    for k, (dfx, dfy) in enumerate(((h, 0.0), (0.0, h))):
        rp = bo.residuals_only(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"] + dfx, P["fy"] + dfy)
        rm = bo.residuals_only(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"] - dfx, P["fy"] - dfy)
        fd = (rp - rm) / (2 * h)
        np.testing.assert_allclose(J[:, :, 9 + k], fd, rtol=1e-7, atol=1e-7 * np.abs(fd).max())


def test_reduced_system_focal_contains_the_camera_system(golden_ba):
    """The (6F+2) reduced system with the focal block among the f-blocks has the constant-focal camera system as its leading
    block (same S, same rhs)."""
    P = _prob(golden_ba, "ring16")
    r, J = bo.residual_jacobian_jets_focal(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    n_cams, n_pts = len(P["cams"]), len(P["pts"])
    R, rhs, g, _ = bo.reduced_system_focal(r, J, P["obs_cam"], P["obs_pt"], n_cams, n_pts, P["cam_const"], 1e-4)
    U, gc, V, gp, W = bo.build_normal_equations(r, J[:, :, :9], P["obs_cam"], P["obs_pt"], n_cams, n_pts, P["cam_const"])
    S, rhs_c, _, _ = bo.schur_reduce(U, gc, V, gp, W, P["obs_cam"], P["obs_pt"], P["cam_const"], 1e-4)
    n6 = S.shape[0]
    np.testing.assert_allclose(R[:n6, :n6], S, rtol=1e-9, atol=1e-9 * np.abs(S).max())
    np.testing.assert_allclose(rhs[:n6], rhs_c, rtol=1e-9, atol=1e-9 * np.abs(rhs_c).max())


def test_lm_focal_recovers_a_perturbed_focal_length():
    P = bo.make_problem(10, 300, 6, 11, noise_px=0.3)
    res = bo.lm_solve_focal(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"] * 1.03, P["fy"] * 0.97)
    assert res["converged"] and res["final_cost"] < 0.05 * res["initial_cost"]
    base = bo.lm_solve(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    assert res["final_cost"] <= 1.05 * base["final_cost"]           # the extra block can only fit better than the true focal ...
    # ... and the gauge (only one camera fixed) lets scale trade against focal length, so the ratio is checked loosely
    assert abs(res["focal"][0] / res["focal"][1] - P["fx"] / P["fy"]) < 0.02


@pytest.mark.parametrize("focal", [False, True])
def test_lm_optimum_vs_scipy_least_squares(golden_ba, focal):
    """An independent solver (MINPACK's Levenberg-Marquardt through scipy.optimize.least_squares, finite-difference
    Jacobian of the plain residual function) started from the same point reaches the same minimum as the restated Ceres
    loop — the oracle's LM rules may differ from Ceres' in path, not in destination.  (Ceres itself is not installed:
    SURVEY.md section 8c.)"""
    scipy_opt = pytest.importorskip("scipy.optimize")
    P = _prob(golden_ba, "small")
    free = ~P["cam_const"].astype(bool)
    n_free, n_pts = int(free.sum()), len(P["pts"])
    fx0, fy0 = (P["fx"] * 1.02, P["fy"] * 0.985) if focal else (P["fx"], P["fy"])

    def unpack(x):
        cams = P["cams"].copy()
        cams[free] = x[:6 * n_free].reshape(-1, 6)
        pts = x[6 * n_free:6 * n_free + 3 * n_pts].reshape(-1, 3)
        f = x[6 * n_free + 3 * n_pts:] if focal else (fx0, fy0)
        return cams, pts, f

    def fun(x):
        cams, pts, f = unpack(x)
        return bo.residuals_only(cams, pts, P["obs_uv"], P["obs_cam"], P["obs_pt"], f[0], f[1]).reshape(-1)

    x0 = np.concatenate([P["cams"][free].reshape(-1), P["pts"].reshape(-1)] + ([np.array([fx0, fy0])] if focal else []))
    sol = scipy_opt.least_squares(fun, x0, method="lm", x_scale="jac", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=200000)
    c_scipy = 0.5 * float((sol.fun ** 2).sum())
    if focal:
        ref = bo.lm_solve_focal(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], fx0, fy0,
                                function_tol=1e-12, max_iters=300)
    else:
        ref = bo.lm_solve(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"],
                          function_tol=1e-12, max_iters=300)
    assert ref["final_cost"] < 0.2 * ref["initial_cost"]
    assert abs(ref["final_cost"] - c_scipy) <= 1e-6 * c_scipy, (ref["final_cost"], c_scipy)
