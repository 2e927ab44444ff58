#!/usr/bin/env python
"""First-light diagnostics for the M-path on a GPU box: never raises, prints as much as possible per call."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import monocularsfm_b200 as m  # noqa: E402
from oracle import match_oracle as mo  # noqa: E402


def compare(tag, idx, dist, d2, oi, od, od2):
    bad_i = np.nonzero(idx[:, 0] != oi[:, 0])[0]
    bad_d0 = np.nonzero(dist[:, 0] != od[:, 0])[0]
    bad_d1 = np.nonzero(dist[:, 1] != od[:, 1])[0]
    print(f"[{tag}] rows={len(idx)} idx0 mismatches={len(bad_i)} dist0 mismatches={len(bad_d0)} dist1 mismatches={len(bad_d1)}")
    for r in list(bad_i[:6]) + list(bad_d1[:6]):
        print(f"   row {r}: got idx={idx[r].tolist()} d2={d2[r].tolist()} dist={dist[r].tolist()} | want idx={oi[r].tolist()} d2={od2[r].tolist()} dist={od[r].tolist()}")
    return len(bad_i) + len(bad_d0) + len(bad_d1) == 0


def main():
    ctx = m.Context(0)
    print("lib version", ctx.lib.msfm_version().decode())
    rng = np.random.default_rng(0)
    ok_all = True
    for (n1, n2) in ((128, 256), (128, 512), (256, 256), (100, 100), (300, 777), (2000, 3000), (8192, 8192)):
        a = rng.integers(0, 256, (n1, 128), dtype=np.uint8)
        b = rng.integers(0, 256, (n2, 128), dtype=np.uint8)
        oi, od, od2 = mo.knn2(a, b)
        for mode in (1, 0):
            try:
                t0 = time.time()
                idx, dist, d2 = ctx.knn2(a, b, mode)
                dt = time.time() - t0
                ok = compare(f"{n1}x{n2} mode{mode} {dt*1e3:.1f}ms", idx, dist, d2, oi, od, od2)
                print("    stats", ctx.match_stats())
                ok_all &= ok
            except Exception:
                traceback.print_exc()
                ok_all = False
                return 1
    print("DIAG", "ALL OK" if ok_all else "MISMATCHES")
    return 0 if ok_all else 2


if __name__ == "__main__":
    sys.exit(main())
