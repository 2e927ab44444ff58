#!/usr/bin/env python
"""Short workloads for ncu captures (never a timing source).
    python tools/prof_run.py match [n_images]     # K1 + post kernels on n_images x 8192 descriptors, all pairs
    python tools/prof_run.py ba                   # one linearisation + one LM solve of BASELINE configs[3]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import monocularsfm_b200 as m  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "match"
    dev = torch.device("cuda", 0)
    ctx = m.Context(0)
    if what == "match":
        n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        descs = bench.make_descriptors_torch(n_img, 8192, 1234, dev)
        for k in range(n_img):
            ctx.upload_dev(k, descs[k].data_ptr(), 8192)
        pairs = bench.all_pairs(n_img)
        cap = len(pairs) * 4096
        d_off = torch.empty(len(pairs) + 1, dtype=torch.int64, device=dev)
        d_mat = torch.empty((cap, 2), dtype=torch.int32, device=dev)
        for _ in range(2):
            tot = ctx.match_pairs_dev(pairs, m.MatchOptions(), d_off.data_ptr(), d_mat.data_ptr(), 0, cap)
        print("matches", tot, ctx.match_stats())
    else:
        P = bench.make_ba_problem(128, 50000, 10.0, 4321)
        ba = ctx.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
        for _ in range(2):
            ba.linearize(1e-4, want_S=False)
        print(ba.solve())
    ctx.close()


if __name__ == "__main__":
    main()
