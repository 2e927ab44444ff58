"""Host-side work partitioning for the one-process-per-GPU model (no compute here).

* ``shard_pairs``       image pairs are independent: interleaved static partition, no collective.
* ``shard_ba_problem``  points (with all their observations) are split across ranks balanced by observation count;
                        cameras are replicated; the partial reduced camera systems are summed with one all-reduce.
"""
from __future__ import annotations

import numpy as np


def shard_pairs(pairs, rank: int, world: int):
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    return pairs[rank::world]


def shard_ba_problem(P: dict, rank: int, world: int) -> dict:
    """Contiguous ranges of points with (nearly) equal observation counts.  Returns a problem dict with the same keys;
    point indices are re-based to the local range (``pt_offset`` gives the global index of local point 0)."""
    obs_pt = np.asarray(P["obs_pt"], np.int64)
    n_pts = len(P["pts"])
    counts = np.bincount(obs_pt, minlength=n_pts)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [int(np.searchsorted(cum, total * k / world, side="left")) for k in range(world + 1)]
    bounds[0], bounds[-1] = 0, n_pts
    lo, hi = bounds[rank], bounds[rank + 1]
    sel = (obs_pt >= lo) & (obs_pt < hi)
    out = dict(P)
    out["pts"] = np.asarray(P["pts"])[lo:hi]
    out["obs_uv"] = np.asarray(P["obs_uv"])[sel]
    out["obs_cam"] = np.asarray(P["obs_cam"])[sel]
    out["obs_pt"] = (obs_pt[sel] - lo).astype(np.int32)
    out["pt_offset"] = lo
    return out
