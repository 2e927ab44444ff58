#!/usr/bin/env python
"""Generate tests/golden/ba_golden.npz from the float64 oracle (oracle/ba_oracle.py) — run in the dev container:

    python tests/golden/gen_ba_golden.py

Ceres is not available anywhere in this image, so these vectors pin the RESTATEMENT (Jet autodiff of the reference's
functor + Ceres' published rotation/Schur/LM formulas), not Ceres itself: B-path parity with Ceres is "unpinned"
(see the oracle's docstring).  The vectors freeze the oracle's outputs so that later edits of the oracle or of the
CUDA path cannot drift unnoticed.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ba_oracle as bo  # noqa: E402


def special_problem():
    """Edge cases of SURVEY §8c: rvec = 0 and |rvec| ~ 1e-9 (Taylor branch), a constant camera, a point with a
    single observation, a long track (> 32 observations), a small-angle camera in the series range."""
    P = bo.make_problem(40, 60, 5, 3)
    cams = P["cams"].copy()
    cams[1, :3] = [1e-9, -2e-9, 0.5e-9]            # theta^2 ~ 5e-18 < DBL_EPSILON -> Taylor branch
    cams[2, :3] = [1e-5, 2e-5, -1e-5]              # tiny but main branch
    cams[3, :3] = [0.05, -0.02, 0.03]              # series range
    P["cams"] = cams
    # rebuild observations: point 0 seen once, point 1 by all 40 cameras, the others as generated
    oc, op = list(P["obs_cam"]), list(P["obs_pt"])
    keep = [(c, p) for c, p in zip(oc, op) if p >= 2]
    obs = [(0, 0)] + [(c, 1) for c in range(40)] + keep
    obs.sort(key=lambda t: (t[1], t[0]))
    P["obs_cam"] = np.array([c for c, _ in obs], np.int32)
    P["obs_pt"] = np.array([p for _, p in obs], np.int32)
    rng = np.random.default_rng(5)
    # put every point in front of every camera that sees it: cameras 0..3 have (nearly) identity rotation -> move them
    # far back along z; regenerate observations from the (modified) geometry plus noise
    cams[1:4, 3:] = [0.0, 0.0, 12.0]
    cams[0, 3:] = [0.0, 0.0, 10.0]
    uv = bo.residuals_only(cams, P["pts"], np.zeros((len(obs), 2)), P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
    P["obs_uv"] = uv + rng.normal(0, 0.5, uv.shape)
    P["cam_const"][:] = 0
    P["cam_const"][0] = 1
    P["cam_const"][7] = 1
    return P


def main():
    store = {}
    probs = {"small": bo.make_problem(8, 200, 6, 0), "special": special_problem(), "ring16": bo.make_problem(16, 600, 8, 1)}
    for name, P in probs.items():
        args = (P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
        r, J = bo.residual_jacobian_jets(P["cams"], P["pts"], *args)
        nc, npnt = len(P["cams"]), len(P["pts"])
        U, gc, V, gp, W = bo.build_normal_equations(r, J, P["obs_cam"], P["obs_pt"], nc, npnt, P["cam_const"])
        S, rhs, Vinv, fmap = bo.schur_reduce(U, gc, V, gp, W, P["obs_cam"], P["obs_pt"], P["cam_const"], 1e-4)
        res = bo.lm_solve(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
        for k, v in P.items():
            store[f"{name}/{k}"] = np.asarray(v)
        store[f"{name}/r"] = r
        store[f"{name}/J"] = J
        store[f"{name}/S"] = S
        store[f"{name}/rhs"] = rhs
        store[f"{name}/gc_free"] = gc[np.asarray(P["cam_const"]) == 0].reshape(-1)
        store[f"{name}/cost"] = np.array(bo.cost_of(r))
        store[f"{name}/lm_costs"] = np.array(res["costs"])
        store[f"{name}/lm_final_cams"] = res["cams"]
        store[f"{name}/lm_final_pts"] = res["pts"]
        store[f"{name}/lm_converged"] = np.array(res["converged"])
        print(name, "obs", len(r), "cost", bo.cost_of(r), "->", res["final_cost"], "iters", res["iterations"],
              "converged", res["converged"], "min depth ok", np.isfinite(r).all())
    path = os.path.join(os.path.dirname(__file__), "ba_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
