#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MonocularSfM hot path (driver contract in the task text).

A "step" is one pass of the matching hot path over one batch of synthetic input: ALL image pairs of
BASELINE.json configs[1] (128 images x 8192 SIFT-like 128-D uint8 descriptors -> 8128 pairs, cross-checked 2-NN
ratio matching).  Metric: descriptor-pairs/s, counting each unordered image pair's n1*n2 descriptor pairs once
(SURVEY.md §8d).  Multi-GPU (torchrun, one rank per GPU): the workload is BASELINE configs[2] — 1329 images, ONE list
of 882 456 pairs — sharded pairs[rank::world] with no data-path collective (descriptors replicated); the total work is the
same for every N > 1 ("scaling": "strong"), `value` = all pairs / the slowest rank's time.  The BA section shards the points
of ONE problem over the ranks (one fp32 + fp64 NCCL all-reduce of the block-sparse system per linearisation).
Further keys: `distributions` (uniform / dense-overlap inputs), `target` (north_star's 1000-image pass), `geometric_verification`
(batched F-matrix RANSAC vs cv2.findFundamentalMat), `ba` / `ba_large` (configs[3] / configs[4] shapes), `roofline.k2`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "descriptor-pairs/s (matching)"
UNIT = "descriptor-pairs/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------------------------- synthetic data
def make_descriptors_torch(n_img, n_desc, seed, device, distribution="S"):
    """Synthetic descriptor sets of SURVEY 8d, generated on the device.
    "S" SIFT-like: |N(0,1)| vectors, L2-normalised to 512, clipped to 255, rounded; image k>0 shares a planted 30 % of
        image 0's descriptors (random slots, +-2 noise) so that the ratio / cross-check lists are non-trivial; plus planted
        duplicates and rows in the float-sqrt collapse range (plant_ties) so that the tie paths run in every pass;
    "D" dense overlap: the same with 90 % planted (most rows pass the ratio test: the worst case of the reverse pass);
    "U" i.i.d. uniform bytes (almost nothing passes the ratio test)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_img, n_desc, 128), dtype=torch.uint8, device=device)
    if distribution == "U":
        for k in range(n_img):
            out[k] = torch.randint(0, 256, (n_desc, 128), generator=g, device=device, dtype=torch.int32).to(torch.uint8)
        return out
    frac = 0.9 if distribution == "D" else 0.3
    base = None
    for k in range(n_img):
        x = torch.randn((n_desc, 128), generator=g, device=device).abs_()
        x = x / x.norm(dim=1, keepdim=True) * 512.0
        x = x.round_().clamp_(0, 255)
        if k == 0:
            base = x.clone()
        else:
            m = int(frac * n_desc)
            dst = torch.randperm(n_desc, generator=g, device=device)[:m]
            src = torch.randperm(n_desc, generator=g, device=device)[:m]
            noise = torch.randint(-2, 3, (m, 128), generator=g, device=device).float()
            x[dst] = (base[src] + noise).clamp_(0, 255)
        plant_ties(x, base, k)
        out[k] = x.to(torch.uint8)
    return out


def plant_ties(x, base, k):
    """Tie coverage of SURVEY 8d in every S / D image: rows 0-3 are image 0's rows 0-3 and rows 4-7 duplicate them, so a query
    that matches one of them finds TWO columns at the best distance (OpenCV keeps the lower index, the ratio test then fails: the
    exact path).  In every second image row 8 is all 255 and rows 9-11 differ from it in one component: matched against an
    image without such rows all their distances are >= 2^22, where distinct integers share one float sqrt (the collapse range);
    against an image that has them they are ties and near-ties at distance 0 / 1 / 2."""
    if len(x) < 16:
        return
    x[0:4] = base[0:4]
    x[4:8] = x[0:4]
    if k % 2 == 0:
        x[8:12] = 255
        x[9, 0] = 254
        x[10, 1] = 254
        x[11, 0] = 253


def make_descriptors_numpy(n_img, n_desc, seed):
    rng = np.random.default_rng(seed)
    out = np.empty((n_img, n_desc, 128), np.uint8)
    base = None
    for k in range(n_img):
        x = np.abs(rng.standard_normal((n_desc, 128), dtype=np.float32))
        x = np.clip(np.rint(x / np.linalg.norm(x, axis=1, keepdims=True) * 512.0), 0, 255)
        if k == 0:
            base = x.copy()
        else:
            m = int(0.3 * n_desc)
            dst = rng.permutation(n_desc)[:m]
            src = rng.permutation(n_desc)[:m]
            x[dst] = np.clip(base[src] + rng.integers(-2, 3, (m, 128)), 0, 255)
        plant_ties(x, base, k)
        out[k] = x.astype(np.uint8)
    return out


def all_pairs(n_img, base_id=0):
    # BruteFeatureMatcher::RunMatching order (FeatureMatching.cpp:110-142): for i, for j < i: (i, j)
    return np.array([(base_id + i, base_id + j) for i in range(n_img) for j in range(i)], np.int32).reshape(-1, 2)


# ----------------------------------------------------------------------------------------------- synthetic BA graph
def _rodrigues_batch(rvec):
    th = np.linalg.norm(rvec, axis=1, keepdims=True)
    k = rvec / np.maximum(th, 1e-300)
    K = np.zeros((len(rvec), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    return np.eye(3)[None] + np.sin(th)[:, :, None] * K + (1 - np.cos(th))[:, :, None] * (K @ K)


def make_ba_problem(n_cams, n_pts, mean_track, seed, noise_px=0.5):
    """Synthetic BA graph of SURVEY.md section 8d (vectorised): cameras on a ring of radius 10 looking inward, camera 0 =
    identity rotation and constant, points uniform in [-3,3]^3, each point seen by its k ~ clip(Geometric(mean_track), 2, .)
    nearest-angle cameras, observations = projection + N(0, 0.5 px), NEU intrinsics, then points / poses perturbed."""
    rng = np.random.default_rng(seed)
    fx = fy = 1449.2752980237
    ang = np.linspace(0, 2 * np.pi, n_cams, endpoint=False)
    centers = np.stack([10 * np.sin(ang), np.zeros(n_cams), -10 * np.cos(ang)], 1)
    rvec = np.stack([np.zeros(n_cams), ang, np.zeros(n_cams)], 1)
    rvec[1:] += rng.normal(0, 0.02, (n_cams - 1, 3))
    R = _rodrigues_batch(rvec)
    tvec = -(R @ centers[:, :, None])[:, :, 0]
    cams = np.concatenate([rvec, tvec], 1)
    pts = rng.uniform(-3, 3, (n_pts, 3))
    k = np.clip(rng.geometric(1.0 / mean_track, n_pts), 2, min(n_cams, 64)).astype(np.int64)
    pang = np.arctan2(pts[:, 0], -pts[:, 2]) % (2 * np.pi)
    c0 = np.rint(pang / (2 * np.pi) * n_cams).astype(np.int64)
    obs_pt = np.repeat(np.arange(n_pts), k)
    within = np.arange(len(obs_pt)) - np.repeat(np.cumsum(k) - k, k)
    off = ((within + 1) // 2) * np.where(within % 2 == 1, 1, -1)          # 0, +1, -1, +2, -2, ...
    obs_cam = (np.repeat(c0, k) + off) % n_cams
    order = np.lexsort((obs_cam, obs_pt))
    obs_cam, obs_pt = obs_cam[order].astype(np.int32), obs_pt[order].astype(np.int32)
    p = (R[obs_cam] @ pts[obs_pt][:, :, None])[:, :, 0] + tvec[obs_cam]
    uv = np.stack([fx * p[:, 0] / p[:, 2], fy * p[:, 1] / p[:, 2]], 1) + rng.normal(0, noise_px, (len(obs_cam), 2))
    cam_const = np.zeros(n_cams, np.uint8)
    cam_const[0] = 1
    pts = pts + rng.normal(0, 0.01, pts.shape)
    cams = cams.copy()
    cams[1:] += rng.normal(0, 0.005, (n_cams - 1, 6))
    return {"cams": cams, "pts": pts, "obs_uv": uv, "obs_cam": obs_cam, "obs_pt": obs_pt, "cam_const": cam_const,
            "fx": fx, "fy": fy}


def ceres_probe():
    """Is the reference's real BA backend (Ceres) on this box?  (BASELINE.md section 3 prefers it over the restatement.)"""
    import ctypes.util
    found = ctypes.util.find_library("ceres")
    ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref")) or os.path.isdir(os.path.join(ROOT, "oracle", "_ref"))
    return {"ceres": found or "absent", "reference_build": "present" if ref else "absent (the reference needs OpenCV C++/Ceres/Eigen headers: DESIGN.md section 5)"}


def bench_ba(ctx, args, rank, world, dist, dev, peaks, n_cams=None, n_pts=None, track=None, cpu_max_reps=20, tag="k2"):
    """BASELINE configs[3] by default: 128 cams / 50k points / ~500k observations (configs[4] shapes when called with
    1329 / 542k / 9.2).  One "BA iteration" = evaluate all residual blocks + Jacobians + Schur-eliminate onto the camera
    system (+ the all-reduce when world > 1).  STRONG scaling: the same problem for every N, points sharded over the
    ranks (monocularsfm_b200.sharding.shard_ba_problem), cameras replicated.  Returns a dict."""
    import torch
    from monocularsfm_b200.sharding import shard_ba_problem
    n_cams = n_cams or args.ba_cams
    n_pts = n_pts or args.ba_pts
    track = track or args.ba_track
    P = make_ba_problem(n_cams, n_pts, track, 4321)
    L = shard_ba_problem(P, rank, world) if world > 1 else P
    t0 = time.perf_counter()
    ba = ctx.ba_create(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["cam_const"], L["fx"], L["fy"])
    create_s = time.perf_counter() - t0
    st = ba.structure()
    n_obs_total = len(P["obs_cam"])
    iters = max(10, args.steps * 4)
    for _ in range(3):
        ba.linearize(1e-4, want_S=False)
    ctx.prof_enable(True)
    ctx.prof_reset()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        ba.linearize(1e-4, want_S=False)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    k_ms_local = prof["ba_schur"]["ms"] / max(1, prof["ba_schur"]["launches"])
    comm_ms_local = prof["ba_comm"]["ms"] / max(1, prof["ba_comm"]["launches"]) if prof["ba_comm"]["launches"] else 0.0
    t = torch.tensor([ms, k_ms_local, comm_ms_local], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, k_ms, comm_ms = float(t[0]), float(t[1]), float(t[2])
    # algorithmic bytes of one launch on this rank (SURVEY 8d, fp64 parameter / observation storage as resident here):
    # 16 B (u, v) + 4 B camera index + 1 B local camera index per observation, 4 B point order + 24 B per point, 184 B per
    # camera, and one write of the reduced system that exists: 144 B per non-empty 6x6 block + the fp64 tail
    alg_bytes = 21.0 * len(L["obs_cam"]) + 28.0 * len(L["pts"]) + 184.0 * len(L["cams"]) + 144.0 * st["n_blocks"] + 8.0 * st["tail_f64"]
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        ent = tj.get(f"{n_cams}x{n_pts}")
        if ent and world == 1:          # the capture is of one GPU holding the whole problem
            traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
    ach = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None
    out = {"workload": f"BA {len(P['cams'])} cams / {len(P['pts'])} points / {n_obs_total} observations, points sharded over {world} GPU(s)",
           "metric": "observations/s per BA iteration (evaluate + Jacobians + Schur reduction" + (" + all-reduce)" if world > 1 else ")"),
           "value": n_obs_total / (ms * 1e-3), "unit": "observations/s", "ms_per_iteration": ms, "scaling": "strong",
           "kernel_ms": k_ms, "allreduce_ms": comm_ms, "allreduce_message_bytes": st["system_bytes"] - 16,
           "structure": st, "create_s": create_s,
           "dtype": "f64 geometry and gradients / f32 6x6 block products and block accumulation",
           "roofline": {"bound": "hbm", "kernel": "fused_linearize_kernel (one launch per linearisation)", "achieved": ach,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (ach / peaks["hbm_gbs"]) if ach else None,
                        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                        "note": "the kernel is bound by issue slots (fp64 geometry + shared-memory compare-and-swap accumulation of "
                                "the 6x6 products), not by HBM: see DESIGN.md B-path"}}
    # full LM solve (the call CeresBundelOptimizer::Optimize makes); the first solve pays one-time library initialisation
    # (cuSOLVER handle, lazy kernel loading), so it is repeated from the same start and the second one is reported
    cams0, pts0 = ba.get_params()
    t0 = time.perf_counter()
    summ = ba.solve()
    first_s = time.perf_counter() - t0
    ba.set_params(cams0, pts0)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    summ = ba.solve()
    out["lm"] = {"iterations": summ["iterations"], "termination": summ["termination"], "initial_cost": summ["initial_cost"],
                 "final_cost": summ["final_cost"], "wall_s": time.perf_counter() - t0, "first_call_wall_s": first_s,
                 "linearize_s": summ["linearize_time_s"], "solve_backsub_eval_s": summ["solve_time_s"],
                 "rmse_px": float(np.sqrt(2 * summ["final_cost"] / max(1, summ["num_residuals"])))}
    ba.close()
    # end to end through the host-buffer C-ABI, as CeresBundelOptimizer::Optimize uses it: structure analysis + H2D of the
    # problem, LM solve, D2H of the parameters
    t0 = time.perf_counter()
    ba2 = ctx.ba_create(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["cam_const"], L["fx"], L["fy"])
    create_s = time.perf_counter() - t0
    up0 = ba2.last_upload()
    s2 = ba2.solve()
    ba2.get_params()
    e2e_s = time.perf_counter() - t0
    out["e2e"] = {"wall_s": e2e_s, "observations_per_s_per_iteration": n_obs_total * s2["iterations"] / e2e_s,
                  "create_s": create_s, "h2d_bytes": up0["h2d_bytes"],
                  "d2h_bytes": int(L["cams"].nbytes + L["pts"].nbytes),
                  "what": "msfm_ba_create (host structure analysis + upload) + msfm_ba_solve + msfm_ba_get_params"}
    # the next Optimize calls on the SAME object (msfm_ba_update; MapBuilder keeps one optimizer, MapBuilder.cpp:92):
    # (a) the same map with the start values again — structure kept, only values travel; (b) a changed map (the last 2 % of the
    # points dropped) — analysed again into the same device arena
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    reused = ba2.update(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["cam_const"], L["fx"], L["fy"])
    upd_s = time.perf_counter() - t0
    up1 = ba2.last_upload()
    s3 = ba2.solve()
    ba2.get_params()
    same_s = time.perf_counter() - t0
    out["e2e"]["next_call_same_map"] = {"wall_s": same_s, "update_s": upd_s, "structure_reused": bool(reused), "h2d_bytes": up1["h2d_bytes"],
                                        "iterations": s3["iterations"], "final_cost_rel_diff": abs(s3["final_cost"] - s2["final_cost"]) / s2["final_cost"]}
    keep_pts = max(1, int(len(L["pts"]) * 0.98))
    keep_obs = int(np.searchsorted(L["obs_pt"], keep_pts, side="left"))
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    reused = ba2.update(L["cams"], L["pts"][:keep_pts], L["obs_uv"][:keep_obs], L["obs_cam"][:keep_obs], L["obs_pt"][:keep_obs], L["cam_const"],
                        L["fx"], L["fy"])
    upd_s = time.perf_counter() - t0
    up2 = ba2.last_upload()
    s4 = ba2.solve()
    ba2.get_params()
    out["e2e"]["next_call_changed_map"] = {"wall_s": time.perf_counter() - t0, "update_s": upd_s, "structure_reused": bool(reused),
                                           "h2d_bytes": up2["h2d_bytes"], "iterations": s4["iterations"], "termination": s4["termination"]}
    ba2.close()
    if rank == 0 and world == 1 and args.cpu_pairs != 0:
        from oracle import ba_oracle as bo
        lib = bo.c_oracle()
        if lib is not None:
            reps, tt = 0, 0.0
            while tt < 5.0 and reps < cpu_max_reps:
                _, _, _, dt = bo.c_linearize(P, 1e-4, lib)
                tt += dt
                reps += 1
            out["cpu_baseline"] = {"value": n_obs_total * reps / tt, "unit": "observations/s", "cores": 1, "kind": "port",
                                   "sample": f"{reps} full evaluate+Schur passes of oracle/ba_oracle.c (float64 Jets + dense Schur, "
                                             "1 thread like the reference's Ceres call, which never sets num_threads)",
                                   **ceres_probe()}
            # the fairer all-cores bound SURVEY 8(d) asks for beside it (NOT what the reference does): the same pass on every
            # host thread, points handed out dynamically, atomic / per-thread accumulation of S
            ncpu = os.cpu_count() or 1
            if ncpu > 1 and hasattr(lib, "ba_oracle_linearize_mt"):
                reps, tt = 0, 0.0
                while tt < 3.0 and reps < max(2, cpu_max_reps):
                    _, _, _, dt = bo.c_linearize(P, 1e-4, lib, threads=ncpu)
                    tt += dt
                    reps += 1
                out["cpu_baseline"]["all_cores"] = {"value": n_obs_total * reps / tt, "unit": "observations/s", "cores": ncpu,
                                                    "sample": f"{reps} passes of the multi-threaded variant of the same port (pthreads)"}
    return out


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.first = 0

    def mark(self):
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        use = self.lines[self.first:] or self.lines[-3:]
        for ln in use:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pairs(descs_np, pairs, distance_ratio=0.8, f32=False):
    """The reference's own CPU path for the given pairs: OpenCV BFMatcher knnMatch both directions
    (FeatureUtils.cpp:160-174) + ratio test + CrossCheck.  Returns seconds."""
    import cv2
    from oracle import match_oracle as mo
    cv2.setNumThreads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    nm = 0
    for (i, j) in pairs:
        a, b = descs_np[i], descs_np[j]
        if f32:
            a, b = a.astype(np.float32), b.astype(np.float32)
        m, d = mo.cv2_match_image_pair(a, b, distance_ratio, -1.0, True, True)
        nm += len(m)
    return time.perf_counter() - t0, nm


def workload_of(args, world):
    """The matching workload of an N-GPU run: BASELINE configs[1] on one GPU, configs[2] (ONE pair list, sharded) on several."""
    n_img = args.images if args.images > 0 else (128 if world == 1 else 1329)
    n = args.ndesc
    P = n_img * (n_img - 1) // 2
    name = (f"match {n_img} images x {n} descriptors, all {P} pairs (cross-check, ratio 0.8)"
            + ("" if world == 1 else f", ONE pair list sharded over {world} GPUs"))
    return n_img, n, P, name


def reference_sample_pairs(args):
    """Image pairs of the CPU sample per step: 32 unless the step count would push the arm beyond a few minutes
    (~0.5 s per pair on 16 cores); never fewer than 8 per step."""
    if args.cpu_pairs > 0:
        return args.cpu_pairs
    return int(max(8, min(32, 300 // max(1, args.steps + min(args.warmup, 1)))))


def run_reference_arm(args, rank, world):
    """--impl reference: OpenCV+glue on the host cores, bounded sample per step (rank 0 only)."""
    if rank != 0:
        return
    n_img, n, P, name = workload_of(args, world)
    sample_pairs = reference_sample_pairs(args)
    descs = make_descriptors_numpy(sample_pairs + 1, n, 1234)
    pairs = [(k + 1, 0) for k in range(sample_pairs)]
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_pairs(descs, pairs[:2])
    t = 0.0
    for _ in range(args.steps):
        dt, _nm = cpu_reference_pairs(descs, pairs)
        t += dt
    val = args.steps * sample_pairs * float(n) * n / t
    cores = os.cpu_count() or 1
    sample = (f"{sample_pairs} of the {P} image pairs per step ({n}x{n} u8 descriptors each), extrapolated linearly; "
              f"cv2 BFMatcher.knnMatch both directions + ratio + CrossCheck, cv2 threads={cores}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": name, "distribution": "SIFT-like, 30% planted correspondences"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             **ceres_probe()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def _two_view(rng, n_in, n_out, noise=0.4):
    """Keypoints of two pinhole views (NEU intrinsics) of random 3-D points: n_in true correspondences with pixel noise, n_out
    random wrong ones; matches (queryIdx, trainIdx) in random order (same generator as tests/test_verify_gpu.py)."""
    K = np.array([[1449.2752980237, 0, 1080.0], [0, 1449.2752980237, 720.0], [0, 0, 1]])
    X = np.c_[rng.uniform(-4, 4, n_in), rng.uniform(-3, 3, n_in), rng.uniform(6, 14, n_in)]
    ang = rng.uniform(0.1, 0.25)
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([-1.5, 0.1, 0.3]) * rng.uniform(0.7, 1.3)
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:]
    x2 = (K @ (R @ X.T + t[:, None])).T
    x2 = x2[:, :2] / x2[:, 2:]
    x1 = x1 + rng.normal(0, noise, x1.shape)
    x2 = x2 + rng.normal(0, noise, x2.shape)
    o1 = np.c_[rng.uniform(0, 2160, n_out), rng.uniform(0, 1440, n_out)]
    o2 = np.c_[rng.uniform(0, 2160, n_out), rng.uniform(0, 1440, n_out)]
    n = n_in + n_out
    perm1, perm2 = rng.permutation(n), rng.permutation(n)
    kp1 = np.zeros((n, 2), np.float32)
    kp2 = np.zeros((n, 2), np.float32)
    kp1[perm1] = np.r_[x1, o1]
    kp2[perm2] = np.r_[x2, o2]
    order = rng.permutation(n)
    return kp1, kp2, np.c_[perm1[order], perm2[order]].astype(np.int32), order < n_in


def bench_dropin_database(n_img=32, n_desc=8192, cpu_pairs=6):
    """SURVEY 8f-1, end to end as a user of the reference runs it: BruteFeatureMatcher(db).RunMatching() of the drop-in C++
    classes (build/host_test) on a SQLite database in the reference's schema — float32 unit-norm descriptor blobs and keypoints
    read from the database, one upload per image, all pairs matched (ratio 0.8, cross-check, max_distance 0.7), F-matrix
    verification of every pair on the device, one matches row per pair written back (FeatureMatching.cpp:10-145).  The wall time is
    that of the whole PROCESS (CUDA context creation and the SQLite I/O included).  Beside it the reference's per-pair CPU cost:
    cv2 BFMatcher both directions on the float32 rows + cv2.findFundamentalMat, on a sample of pairs."""
    import sqlite3
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "build", "host_test")
    if not os.path.exists(exe):
        return {"skipped": "build/host_test is missing (make)"}
    rng = np.random.default_rng(2024)
    K = np.array([[1449.2752980237, 0, 1080.0], [0, 1449.2752980237, 720.0], [0, 0, 1]])
    X = np.c_[rng.uniform(-4, 4, n_desc), rng.uniform(-3, 3, n_desc), rng.uniform(6, 14, n_desc)]      # one 3-D point per base row

    def sift_like(n):
        x = np.abs(rng.standard_normal((n, 128), dtype=np.float32))
        return np.clip(np.rint(x / np.linalg.norm(x, axis=1, keepdims=True) * 512.0), 0, 255)
    base = sift_like(n_desc)
    tmp = tempfile.mkdtemp(prefix="msfm_dropin_")
    db = os.path.join(tmp, "bench.db")
    con = sqlite3.connect(db)
    con.executescript("""
        CREATE TABLE images(image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE);
        CREATE TABLE keypoints(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE colors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE descriptors(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
        CREATE TABLE matches(pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
    """)
    descs, kps = [], []
    for k in range(n_img):
        ang = 0.02 * k
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        t = np.array([-0.25 * k, 0.02 * k, 0.03 * k])
        d = sift_like(n_desc)
        kp = np.zeros((n_desc, 4), np.float32)
        kp[:, 0] = rng.uniform(0, 2160, n_desc)
        kp[:, 1] = rng.uniform(0, 1440, n_desc)
        kp[:, 2] = rng.permutation(n_desc).astype(np.float32) + 1.0
        m = int(0.3 * n_desc)
        dst, src = rng.permutation(n_desc)[:m], rng.permutation(n_desc)[:m]
        d[dst] = np.clip(base[src] + rng.integers(-2, 3, (m, 128)), 0, 255)
        x = (K @ (R @ X[src].T + t[:, None])).T
        kp[dst, :2] = (x[:, :2] / x[:, 2:] + rng.normal(0, 0.4, (m, 2))).astype(np.float32)      # planted rows see their 3-D point
        con.execute("insert into images(image_id, name) values(?, ?)", (k, f"img{k}.jpg"))
        con.execute("insert into keypoints values(?,?,?,?)", (k, n_desc, 4, kp.tobytes()))
        con.execute("insert into descriptors values(?,?,?,?)", (k, n_desc, 128, (d / 512.0).astype(np.float32).tobytes()))
        descs.append((d / 512.0).astype(np.float32))
        kps.append(kp[:, :2].copy())
    con.commit()
    con.close()
    n_pairs = n_img * (n_img - 1) // 2
    t0 = time.perf_counter()
    # the reference's default max_pairs_size (100): five device calls / transactions for the 496 pairs.  (ONE batch of all 496
    # pairs measured slower here, 1.5 s instead of ~0.5 s inside RunMatching: the first call then allocates the scratch of a
    # 65 k-unit batch, 1.5 GB, for a single use.)
    run = subprocess.run([exe, "match", db, "0", "quiet", "100"], capture_output=True, text=True, timeout=600)
    wall = time.perf_counter() - t0
    inner = [l for l in run.stdout.splitlines() if l.startswith("RunMatching:")]
    out = {"workload": f"BruteFeatureMatcher::RunMatching on a reference-schema SQLite database: {n_img} images x {n_desc} float32 descriptors, "
                       f"all {n_pairs} pairs, geometric verification on",
           "wall_s": wall, "returncode": run.returncode, "pairs_per_s": n_pairs / wall,
           "descriptor_pairs_per_s": float(n_pairs) * n_desc * n_desc / wall,
           "run_matching_s": float(inner[-1].split()[1]) if inner else None, "max_pairs_size": 100,
           "what": "wall_s = whole process (program start, CUDA context); run_matching_s = RunMatching() alone: SQLite reads of descriptors / "
                   "keypoints, uploads, matching, verification, SQLite writes (the CUDA context is created inside it)"}
    if run.returncode != 0:
        out["error"] = (run.stderr + run.stdout)[-500:]
        return out
    con = sqlite3.connect(db)
    rows = con.execute("select count(*), sum(rows) from matches").fetchone()
    con.close()
    out["match_rows_written"], out["verified_matches"] = int(rows[0]), int(rows[1] or 0)
    # the reference's cost per pair on this box's host cores, on a sample
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        bf = cv2.BFMatcher(cv2.NORM_L2)
        t0 = time.perf_counter()
        for q in range(cpu_pairs):
            i, j = q + 1, q
            m12 = bf.knnMatch(descs[i], descs[j], k=2)
            m21 = bf.knnMatch(descs[j], descs[i], k=2)
            good = [a for a, b in m12 if a.distance < 0.8 * b.distance]
            back = {a.queryIdx: a.trainIdx for a, b in m21 if a.distance < 0.8 * b.distance}
            cross = [a for a in good if back.get(a.trainIdx, 0) == a.queryIdx and a.distance <= 0.7]
            if len(cross) >= 8:
                cv2.findFundamentalMat(kps[i][[a.queryIdx for a in cross]], kps[j][[a.trainIdx for a in cross]], cv2.FM_RANSAC, 3.0, 0.99)
        per_pair = (time.perf_counter() - t0) / max(1, cpu_pairs)
        out["cpu"] = {"s_per_pair": per_pair, "extrapolated_wall_s": per_pair * n_pairs, "sample_pairs": cpu_pairs, "cores": os.cpu_count(),
                      "kind": "reference (cv2 BFMatcher on the float32 rows, both directions + ratio + cross-check + distance filter, "
                              "cv2.findFundamentalMat), without its SQLite I/O"}
        out["speedup_vs_cpu_extrapolated"] = per_pair * n_pairs / wall
    except Exception as ex:      # cv2 missing: the GPU figure stands alone
        out["cpu"] = {"error": repr(ex)}
    try:
        os.remove(db)
        os.rmdir(tmp)
    except OSError:
        pass
    return out


def bench_verify_geometric(ctx, n_pairs=8128, n_scenes=64, n_in=534, n_out=229, cpu_pairs=64):
    """FeatureUtils::FilterMatches (cv::findFundamentalMat FM_RANSAC 3.0 / 0.99, FeatureUtils.cpp:176-206) over the
    matches of every pair of the headline workload's size (8128 pairs x ~763 matches, 70 % true correspondences), batched on
    the device through the host-pointer C-ABI call vs cv2.findFundamentalMat per pair on the host cores."""
    import cv2
    rng = np.random.default_rng(99)
    scenes = [_two_view(rng, n_in, n_out) for _ in range(n_scenes)]
    for k, (kp1, kp2, _m, _t) in enumerate(scenes):
        ctx.upload_keypoints(200000 + 2 * k, kp1)
        ctx.upload_keypoints(200001 + 2 * k, kp2)
    pairs = [(200000 + 2 * (i % n_scenes), 200001 + 2 * (i % n_scenes)) for i in range(n_pairs)]
    per = n_in + n_out
    offs = np.arange(n_pairs + 1, dtype=np.int64) * per
    matches = np.concatenate([scenes[i % n_scenes][2] for i in range(n_pairs)])
    ctx.verify_pairs(pairs[:64], offs[:65], matches[:64 * per])                      # warm-up
    ctx.prof_enable(True)
    ctx.prof_reset()
    t0 = time.perf_counter()
    mask, counts = ctx.verify_pairs(pairs, offs, matches)
    wall = time.perf_counter() - t0
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    kern_ms = prof.get("verify", {}).get("ms", 0.0)
    # host reference on a sample: the call the reference makes, all cv2 threads
    t0 = time.perf_counter()
    agree, kept_of_cv, cv_in = [], 0, 0
    n_cpu = min(cpu_pairs, n_scenes)
    for k in range(n_cpu):
        kp1, kp2, m, _truth = scenes[k]
        _F, cm = cv2.findFundamentalMat(kp1[m[:, 0]], kp2[m[:, 1]], cv2.FM_RANSAC, 3.0, 0.99)
        cm = cm.ravel().astype(bool) if cm is not None else np.zeros(per, bool)
        got = mask[k * per:(k + 1) * per]
        agree.append(float((got == cm).mean()))
        kept_of_cv += int((got & cm).sum())
        cv_in += int(cm.sum())
    cpu_wall = time.perf_counter() - t0
    truth_all = np.concatenate([scenes[i % n_scenes][3] for i in range(n_pairs)])
    return {"workload": f"F-matrix RANSAC (3.0 px, 0.99) over {n_pairs} pairs x {per} matches ({n_in} true + {n_out} wrong), "
                        f"{n_scenes} distinct two-view scenes",
            "pairs_per_s": n_pairs / wall, "wall_s": wall, "kernel_ms": kern_ms,
            "what": "msfm_verify_pairs with HOST match lists: H2D of the matches, one CTA per pair, D2H of the masks",
            "h2d_bytes": int(matches.nbytes + offs.nbytes), "d2h_bytes": int(mask.size + counts.nbytes),
            "cpu": {"pairs_per_s": n_cpu / cpu_wall, "sample_pairs": n_cpu, "cores": os.cpu_count() or 1,
                    "kind": "reference (cv2.findFundamentalMat per pair)"},
            "parity": {"min_agreement_with_cv2": min(agree), "mean_agreement_with_cv2": sum(agree) / len(agree),
                       "cv2_inliers_kept": kept_of_cv / max(1, cv_in),
                       "true_correspondences_kept": float(mask[truth_all].mean()),
                       "wrong_matches_kept": float(mask[~truth_all].mean())}}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def timed_resident(ctx, torch, dev, dist, pairs, opt, bufs, steps, warmup):
    """`steps` passes over `pairs` with descriptors resident; returns (ms total on the library stream, matches of a pass)."""
    d_off, d_mat, d_dst, capacity = bufs
    total = 0
    for _ in range(warmup):
        total = ctx.match_pairs_dev(pairs, opt, d_off.data_ptr(), d_mat.data_ptr(), d_dst.data_ptr(), capacity)
    ctx.sync()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        total = ctx.match_pairs_dev(pairs, opt, d_off.data_ptr(), d_mat.data_ptr(), d_dst.data_ptr(), capacity)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), total


def run_ours(args, rank, world, local_rank):
    import torch
    import monocularsfm_b200 as m
    from monocularsfm_b200.sharding import shard_pairs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    n_img, n, P_all, wl_name = workload_of(args, world)
    ctx = m.Context(local_rank)
    if dist:
        # the library's own NCCL communicator (all-reduce of the reduced camera system); id travels over the
        # torch.distributed store
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    # every rank holds ALL descriptors (1329 x 1 MiB replicated) and its share of the ONE pair list
    if args.cpu_data:      # profiling runs: keep torch's data-generation kernels out of the ncu launch list
        descs = torch.from_numpy(make_descriptors_numpy(n_img, n, 1234)).to(dev)
    else:
        descs = make_descriptors_torch(n_img, n, 1234, dev)              # [n_img, n, 128] u8, resident in HBM
    torch.cuda.synchronize()
    pairs_all = all_pairs(n_img)
    pairs = np.ascontiguousarray(shard_pairs(pairs_all, rank, world)) if world > 1 else pairs_all
    P = len(pairs)
    opt = m.MatchOptions(0.8, -1.0, True, True)
    per_pair_cap = 8192 if P <= 20000 else 1280       # at most one match per query row; the big pair lists hold ~800 per pair
    capacity = int(P) * per_pair_cap + 65536
    d_off = torch.empty(P + 1, dtype=torch.int64, device=dev)
    d_mat = torch.empty((capacity, 2), dtype=torch.int32, device=dev)
    d_dst = torch.empty(capacity, dtype=torch.float32, device=dev)
    bufs = (d_off, d_mat, d_dst, capacity)

    # ---------------- device-resident throughput ("value")
    for k in range(n_img):
        ctx.upload_dev(k, descs[k].data_ptr(), n)
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a moment to come up: start before the warm-up steps
    timed_resident(ctx, torch, dev, None, pairs, opt, bufs, 0, args.warmup)
    sampler.mark()                       # only samples taken from here on (the timed region) are reported
    ctx.prof_enable(True)
    ctx.prof_reset()
    launches0 = ctx.launch_count
    ms_total, total = timed_resident(ctx, torch, dev, dist, pairs, opt, bufs, args.steps, 0)
    if dist:
        dist.barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    stats = ctx.match_stats()

    # ---------------- one sampled pair per step checked against the oracle (outside the timed region)
    verify = None
    if rank == 0 and not args.no_verify:
        from oracle import match_oracle as mo
        rng = np.random.default_rng(99)
        off_h = d_off.cpu().numpy()
        ok, checked = True, []
        for p in rng.choice(P, size=min(P, max(1, args.steps)), replace=False):
            i, j = int(pairs[p][0]), int(pairs[p][1])
            got = d_mat[int(off_h[p]):int(off_h[p + 1])].cpu().numpy()
            gd = d_dst[int(off_h[p]):int(off_h[p + 1])].cpu().numpy()
            em, ed = mo.cv2_match_image_pair(descs[i].cpu().numpy(), descs[j].cpu().numpy(), 0.8, -1.0, True, True)
            same = bool(np.array_equal(got, em) and np.array_equal(gd, ed))
            ok &= same
            checked.append([i, j, int(len(em)), same])
        verify = {"pairs_checked_vs_cv2": checked, "bit_exact": ok}

    # ---------------- end to end through the host-buffer C-ABI call ("e2e"), same number of steps
    host_descs = torch.empty((n_img, n, 128), dtype=torch.uint8, pin_memory=True)
    host_descs.copy_(descs)
    host_np = host_descs.numpy()
    h_off = np.zeros(P + 1, np.int64)
    cap_h = int(total * 1.25) + 1024
    h_mat = torch.empty((cap_h, 2), dtype=torch.int32, pin_memory=True).numpy()
    h_dst = torch.empty(cap_h, dtype=torch.float32, pin_memory=True).numpy()
    import ctypes as C

    def step_e2e():
        for k in range(n_img):
            ctx.upload(k, host_np[k])
        tot = C.c_int64(0)
        rc = ctx.lib.msfm_match_pairs(ctx.h, pairs.ctypes.data_as(C.c_void_p), P, C.byref(opt),
                                      h_off.ctypes.data_as(C.c_void_p), h_mat.ctypes.data_as(C.c_void_p),
                                      h_dst.ctypes.data_as(C.c_void_p), cap_h, C.byref(tot))
        ctx._check(rc)
        return int(tot.value)

    e2e_steps = args.steps
    step_e2e()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tot_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = n_img * n * 128 + pairs.nbytes
    d2h = (P + 1) * 8 + tot_e2e * 12 + 64
    del host_descs, h_mat, h_dst

    # ---------------- reduce over ranks (max time); per-rank times for the load imbalance
    times = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    per_rank = [ms_total / args.steps]
    if dist:
        gathered = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, times)
        per_rank = [float(g[0]) / args.steps for g in gathered]
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total_max, e2e_ms_max = float(times[0]), float(times[1])
    work_per_step = float(P_all) * n * n                  # descriptor pairs of the WHOLE job, each unordered image pair once
    value = work_per_step * args.steps / (ms_total_max * 1e-3)
    e2e_value = work_per_step * e2e_steps / (e2e_ms_max * 1e-3)

    # ---------------- other distributions and the north_star target size (N = 1 only; descriptors resident, short runs)
    dists_out, target = None, None
    if world == 1 and not args.no_extra:
        dists_out = {"S": {"value": value, "ms_per_step": ms_total_max / args.steps, "matches_per_step": int(total),
                           "what": "SIFT-like, 30% planted (the headline)"}}
        for tag, what in (("U", "i.i.d. uniform bytes"), ("D", "SIFT-like, 90% planted (dense overlap: most rows pass the ratio test)")):
            dd = make_descriptors_torch(n_img, n, 4242, dev, tag)
            for k in range(n_img):
                ctx.upload_dev(k, dd[k].data_ptr(), n)
            ms_d, tot_d = timed_resident(ctx, torch, dev, None, pairs, opt, bufs, 2, 1)
            dists_out[tag] = {"value": work_per_step * 2 / (ms_d * 1e-3), "ms_per_step": ms_d / 2, "matches_per_step": int(tot_d),
                              "rescans": ctx.match_stats()["rescans"], "what": what}
            del dd
    if world == 1 and not args.no_target:
        # north_star's target: 1000 images x 8192, all 499 500 pairs, one pass, wall clock (descriptors resident)
        del d_mat, d_dst, d_off, bufs
        torch.cuda.empty_cache()
        n_t = args.target_images
        dt_ = make_descriptors_torch(n_t, n, 777, dev)
        ctx.release_all()
        for k in range(n_t):
            ctx.upload_dev(k, dt_[k].data_ptr(), n)
        pt = all_pairs(n_t)
        cap_t = len(pt) * 1280 + 65536
        t_off = torch.empty(len(pt) + 1, dtype=torch.int64, device=dev)
        t_mat = torch.empty((cap_t, 2), dtype=torch.int32, device=dev)
        ctx.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot_t = ctx.match_pairs_dev(pt, opt, t_off.data_ptr(), t_mat.data_ptr(), 0, cap_t)
        ctx.sync()
        wall = time.perf_counter() - t0
        target = {"workload": f"match {n_t} images x {n} descriptors, all {len(pt)} pairs, one pass", "wall_s": wall,
                  "value": float(len(pt)) * n * n / wall, "matches": int(tot_t)}
        del t_mat, t_off, dt_
        ctx.release_all()
        torch.cuda.empty_cache()

    geo = None
    if world == 1 and not args.no_extra:
        try:
            geo = bench_verify_geometric(ctx)
        except Exception as ex:
            geo = {"error": repr(ex)}
    dropin = None
    if world == 1 and not args.no_extra:
        try:
            dropin = bench_dropin_database()
        except Exception as ex:
            dropin = {"error": repr(ex)}
    peaks = load_peaks()
    ba_out = None
    ba_large = None
    if not args.no_ba:
        try:
            ba_out = bench_ba(ctx, args, rank, world, dist, dev, peaks)
        except Exception as ex:                      # the matching headline must survive a BA failure
            ba_out = {"error": repr(ex)}
        if not args.no_ba_large:
            # the shapes north_star quotes its BA target on (BASELINE configs[4]: 1329 cams / 542k points / ~5M observations);
            # the CPU port runs a single pass of it (~6 s, N = 1 only)
            try:
                ba_large = bench_ba(ctx, args, rank, world, dist, dev, peaks, 1329, 542000, 9.2, cpu_max_reps=1)
            except Exception as ex:
                ba_large = {"error": repr(ex)}
    if rank == 0:
        k1 = prof["match_tile"]
        k1_avg_s = (k1["ms"] / max(1, k1["launches"])) * 1e-3
        steps_launches = max(1, k1["launches"])
        # algorithmic int8 ops of one K1 launch on this rank: 256 per descriptor pair, each unordered image pair once
        # (SURVEY 8d); the kernel executes more on the tensor cores (forward + reverse pass for the cross-check).
        alg_ops_per_launch = float(P) * n * n * 256.0 * args.steps / steps_launches
        achieved = alg_ops_per_launch / k1_avg_s / 1e12 if k1_avg_s > 0 else 0.0
        peak_int8 = 2.0 * peaks["bf16_tflops_sustained"]
        # DRAM traffic of one K1 launch of this workload, from the committed `ncu --set full` capture (bytes; null if absent)
        traffic, traffic_src, ncu_ops_pct = None, None, None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
            ncu_ops_pct = tj.get("tensor_pipe_ops_pct_of_peak_ncu")
        step_frac = (float(P) * n * n * 256.0 * args.steps / (ms_total * 1e-3) / 1e12) / peak_int8 if ms_total else None
        roofline = {"bound": "tensor", "kernel": "match_pair_kernel", "achieved": achieved, "peak": peak_int8,
                    "unit": "TOP/s", "ops": "integer: u8 x u8 -> s32 multiply-accumulates on the tensor cores, 2 ops each",
                    "frac": achieved / peak_int8 if peak_int8 else None, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_note": f"2 x cuBLAS bf16 sustained ({peaks['bf16_tflops_sustained']} TF/s, {peaks['source']} "
                                 "MEASURED_PEAKS.json): int8 tcgen05 runs at twice the bf16 rate; no int8 GEMM peak is measured",
                    # what the tensor pipe itself did in the committed ncu capture of a bench-sized launch: executed int8 ops
                    # (forward + reverse pass, 160 K bytes instead of 128, padded columns) as % of the hardware peak of
                    # 16384 ops/clk/SM — a number taken under the profiler, quoted as evidence, not as a bench value
                    "tensor_pipe_executed_ops_pct_of_hw_peak_ncu": ncu_ops_pct,
                    "hw_peak_note": "tcgen05 kind::i8 microbenchmarked at 8192 MAC/clk/SM = 4.77 POP/s at 1965 MHz "
                                    "(profiles/r01_microbench.log); achieved/4770 = %.3f" % (achieved / 4770.0),
                    "avg_launch_ms": k1_avg_s * 1e3, "launches_timed": k1["launches"],
                    "kernel_share_of_step": k1["ms"] / ms_total if ms_total else None,
                    "step_level_frac": step_frac,
                    # the INT32-ALU bound of the epilogue: per 128 x 256 accumulator tile of a CTA one VIMNMX3 per
                    # two elements = 512 warp instructions, 3.7 cycles each per sub-partition with fresh operands
                    # (profiles/r01b_microbench_pair.log) over 4 sub-partitions, against the 640 tensor cycles of the tile's
                    # five MMAs: the reduction fits under the tensor work only because each warpgroup has two tile times per tile
                    "epilogue_alu_bound": {"vimnmx3_warp_instructions_per_tile": 512, "cycles_per_instruction_per_subpartition": 3.7,
                                           "alu_cycles_per_tile": 512 / 4 * 3.7, "tensor_cycles_per_tile": 640,
                                           "alu_over_tensor": 512 / 4 * 3.7 / 640}}
        if ba_large and "roofline" in ba_large:
            roofline["k2"] = dict(ba_large["roofline"], workload=ba_large["workload"], observations_per_s=ba_large["value"],
                                  ms_per_iteration=ba_large["ms_per_iteration"], allreduce_ms=ba_large["allreduce_ms"])
        # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
        cpu = None
        if world == 1 and args.cpu_pairs != 0:
            sample_pairs = args.cpu_pairs if args.cpu_pairs > 0 else 32
            sub = descs[: sample_pairs + 1].cpu().numpy()
            pl = [(k + 1, 0) for k in range(sample_pairs)]
            t_u8, _ = cpu_reference_pairs(sub, pl)
            t_f32, _ = cpu_reference_pairs(sub, pl[: max(1, sample_pairs // 2)], f32=True)
            cores = os.cpu_count() or 1
            cpu = {"value": sample_pairs * float(n) * n / t_u8, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{sample_pairs} of {P_all} image pairs ({n}x{n} u8), extrapolated linearly; OpenCV BFMatcher.knnMatch both "
                             f"directions + ratio + CrossCheck via cv2 (the routine FeatureUtils.cpp:146-149 calls), cv2 threads={cores}",
                   "f32_value": max(1, sample_pairs // 2) * float(n) * n / t_f32,
                   "f32_note": "same pairs as float32 (what the reference's DB really stores; OpenCV's f32 path is faster)",
                   **ceres_probe()}
            if ba_large and "cpu_baseline" in ba_large:
                cpu["ba"] = dict(ba_large["cpu_baseline"], workload=ba_large["workload"],
                                 gpu_over_cpu_per_iteration=ba_large["value"] / ba_large["cpu_baseline"]["value"])
            if target:
                target["cpu_extrapolated_s"] = {"u8": target["value"] * target["wall_s"] / cpu["value"],
                                                "f32": target["value"] * target["wall_s"] / cpu["f32_value"]}
                target["speedup_vs_cpu"] = {"u8": target["cpu_extrapolated_s"]["u8"] / target["wall_s"],
                                            "f32": target["cpu_extrapolated_s"]["f32"] / target["wall_s"]}
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "ms_per_step": e2e_ms_max / e2e_steps}
        if ba_large and "e2e" in ba_large:
            e2e["ba"] = dict(ba_large["e2e"], workload=ba_large["workload"])
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": wl_name, "distribution": "SIFT-like, 30% planted correspondences",
                           "l2": "working set per step (descriptors + ~2.7 GB row scratch per batch) exceeds the 126 MB L2",
                           "pairs_this_rank": int(P), "matches_this_rank_per_step": int(total),
                           "per_rank_ms_per_step": per_rank,
                           "load_imbalance": (max(per_rank) / (sum(per_rank) / len(per_rank))) if per_rank else None},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
                "roofline": roofline, "cpu_baseline": cpu, "verify": verify,
                "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]},
                "match_stats": stats, "distributions": dists_out, "target": target,
                "geometric_verification": geo, "dropin_database": dropin, "ba": ba_out, "ba_large": ba_large}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=0, help="0: BASELINE configs[1] (128) on one GPU, configs[2] (1329) on several")
    ap.add_argument("--ndesc", type=int, default=8192)
    ap.add_argument("--cpu-pairs", type=int, default=-1, help="image pairs in the CPU-baseline sample (-1: 32; 0: skip)")
    ap.add_argument("--no-ba", action="store_true", help="skip the BA measurements")
    ap.add_argument("--no-ba-large", action="store_true", help="skip the configs[4]-sized BA measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the U / dense-overlap distributions (N=1)")
    ap.add_argument("--no-target", action="store_true", help="skip the 1000-image north_star target pass (N=1)")
    ap.add_argument("--no-verify", action="store_true", help="skip the per-step sampled-pair check against cv2")
    ap.add_argument("--target-images", type=int, default=1000)
    ap.add_argument("--cpu-data", action="store_true", help="generate the synthetic descriptors with numpy (ncu launch lists)")
    ap.add_argument("--ba-cams", type=int, default=128)
    ap.add_argument("--ba-pts", type=int, default=50000)
    ap.add_argument("--ba-track", type=float, default=10.0)
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
