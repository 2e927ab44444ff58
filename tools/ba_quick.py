#!/usr/bin/env python
"""Quick device timing of the B-path at the two BASELINE shapes (development aid, not a bench line):
    python tools/ba_quick.py [small|large|both] [iters]
Prints the structure the device built, ms per linearisation (CUDA events on the library stream) and one LM solve."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import monocularsfm_b200 as m  # noqa: E402


def run(ctx, name, n_cams, n_pts, track, iters):
    t0 = time.perf_counter()
    P = bench.make_ba_problem(n_cams, n_pts, track, 4321)
    t1 = time.perf_counter()
    ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
    t2 = time.perf_counter()
    st = ba.structure()
    print(f"{name}: {len(P['obs_cam'])} obs, generate {t1 - t0:.2f}s, create {t2 - t1:.3f}s, structure {st}", flush=True)
    for _ in range(3):
        ba.linearize(1e-4, want_S=False)
    stream = torch.cuda.ExternalStream(ctx.stream)
    ctx.prof_enable(True)
    ctx.prof_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(iters):
        ba.linearize(1e-4, want_S=False)
    e1.record(stream)
    torch.cuda.synchronize()
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    ms = e0.elapsed_time(e1) / iters
    k = prof["ba_schur"]
    print(f"{name}: {ms:.3f} ms per linearize call, kernel {k['ms'] / max(1, k['launches']):.3f} ms, "
          f"{len(P['obs_cam']) / ms / 1e6:.1f} G obs/s... = {len(P['obs_cam']) / (ms * 1e-3):.3e} obs/s", flush=True)
    t0 = time.perf_counter()
    s = ba.solve()
    print(f"{name}: solve {time.perf_counter() - t0:.3f}s", s, flush=True)
    cams0, pts0 = None, None
    t0 = time.perf_counter()
    s = ba.solve()          # from the optimum: one or two iterations; shows the per-iteration cost without first-call effects
    print(f"{name}: second solve {time.perf_counter() - t0:.3f}s iterations {s['iterations']} solver {ba.solver_info()}", flush=True)
    ba.set_params(P["cams"], P["pts"])
    ctx.prof_enable(True)
    ctx.prof_reset()
    t0 = time.perf_counter()
    s = ba.solve()
    wall = time.perf_counter() - t0
    prof = ctx.prof_read()
    ctx.prof_enable(False)
    print(f"{name}: full solve again {wall:.4f}s iterations {s['iterations']} final {s['final_cost']:.6e} "
          + " ".join(f"{k}={v['ms']:.2f}ms/{v['launches']}" for k, v in prof.items() if v['launches']), flush=True)
    ba.close()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "both"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ctx = m.Context(0)
    if what in ("small", "both"):
        run(ctx, "configs[3]", 128, 50000, 10.0, iters)
    if what in ("large", "both"):
        run(ctx, "configs[4]", 1329, 542000, 9.2, iters)
    ctx.close()


if __name__ == "__main__":
    main()
