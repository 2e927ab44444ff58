#!/usr/bin/env python
"""2..8-GPU functional check (launch with torchrun, one rank per GPU):
   - the library's NCCL communicator (dlopen'ed ncclAllReduce) sums / maxes correctly,
   - a point-sharded BA solve with the all-reduced reduced camera system reproduces the single-GPU solve,
   - msfm_ba_update takes the same path on every rank (values only / re-analysis) and reproduces a fresh problem,
   - pair-sharded matching reproduces the single-GPU match lists.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import monocularsfm_b200 as m  # noqa: E402
from monocularsfm_b200.sharding import shard_ba_problem, shard_pairs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = m.Context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    ok = True
    # ---- raw all-reduce
    t = torch.full((1000,), float(rank + 1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ctx.comm_allreduce_f64(t.data_ptr(), t.numel(), 0)
    ctx.sync()
    ok &= bool((t == world * (world + 1) / 2).all())
    t = torch.full((8,), float(rank), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ctx.comm_allreduce_f64(t.data_ptr(), t.numel(), 1)
    ctx.sync()
    ok &= bool((t == world - 1).all())
    if rank == 0:
        print("allreduce sum/max:", ok, flush=True)
    # ---- BA
    P = bench.make_ba_problem(64, 20000, 8.0, 77)
    L = shard_ba_problem(P, rank, world)
    ba = ctx.ba_create(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["cam_const"], L["fx"], L["fy"])
    S, rhs, gc, cost = ba.linearize(1e-4)
    s = ba.solve()
    cams_multi, _ = ba.get_params()
    # msfm_ba_update is collective: (a) same pattern on every rank -> values only; (b) ONE rank's shard changes -> every rank
    # re-analyses (the block structure is merged across ranks); the result must equal a freshly created sharded problem
    reused_a = ba.update(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["cam_const"], L["fx"], L["fy"])
    s_a = ba.solve()
    P2 = bench.make_ba_problem(64, 20000, 8.0, 78)
    L2 = dict(shard_ba_problem(P2, rank, world), cams=L["cams"], cam_const=L["cam_const"]) if rank == 0 else L      # only rank 0's points differ
    reused_b = ba.update(L2["cams"], L2["pts"], L2["obs_uv"], L2["obs_cam"], L2["obs_pt"], L2["cam_const"], L2["fx"], L2["fy"])
    S_b, rhs_b, _, _ = ba.linearize(1e-4)
    fresh = ctx.ba_create(L2["cams"], L2["pts"], L2["obs_uv"], L2["obs_cam"], L2["obs_pt"], L2["cam_const"], L2["fx"], L2["fy"])
    S_f, rhs_f, _, _ = fresh.linearize(1e-4)
    fresh.close()
    e_upd = max(np.abs(S_b - S_f).max() / np.abs(S_f).max(), np.abs(rhs_b - rhs_f).max() / np.abs(rhs_f).max())
    ok_upd = bool(reused_a) and not reused_b and abs(s_a["final_cost"] - s["final_cost"]) <= 1e-7 * s["final_cost"] and e_upd < 1e-6
    if rank == 0:
        print(f"BA update (collective): same pattern reused {reused_a}, changed on one rank reused {reused_b}, "
              f"re-analysed vs fresh {e_upd:.2e}: {ok_upd}", flush=True)
    ok &= ok_upd
    ba.close()
    if rank == 0:
        solo = m.Context(local)                       # no communicator: the whole problem on one GPU
        b1 = solo.ba_create(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["cam_const"], P["fx"], P["fy"])
        S1, rhs1, gc1, cost1 = b1.linearize(1e-4)
        s1 = b1.solve()
        cams_solo, _ = b1.get_params()
        eS = np.abs(S - S1).max() / np.abs(S1).max()
        er = np.abs(rhs - rhs1).max() / np.abs(rhs1).max()
        ec = abs(s["final_cost"] - s1["final_cost"]) / s1["final_cost"]
        ep = np.abs(cams_multi - cams_solo).max()
        print(f"BA sharded vs solo: S {eS:.2e} rhs {er:.2e} cost {abs(cost - cost1) / cost1:.2e} final {ec:.2e} "
              f"cams {ep:.2e} iters {s['iterations']}/{s1['iterations']} nres {s['num_residuals']}/{s1['num_residuals']}", flush=True)
        # the 6x6 blocks are accumulated in fp32 (shared-memory tiles, fp32 reductions, fp32 all-reduce): 1e-7-level differences
        ok &= eS < 1e-6 and er < 1e-6 and ec < 1e-7 and ep < 1e-5 and s["num_residuals"] == s1["num_residuals"] and s["termination"] == 0
        b1.close()
        solo.close()
    # ---- matching
    n_img = 6
    descs = bench.make_descriptors_numpy(n_img, 2048, 5)
    for k in range(n_img):
        ctx.upload(k, descs[k])
    pairs = bench.all_pairs(n_img)
    mine = shard_pairs(pairs, rank, world)
    off, mt, d = ctx.match_pairs(mine, m.MatchOptions())
    counts = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
    counts[rank::world] = torch.from_numpy(np.diff(off)).to(dev)
    dist.all_reduce(counts)
    if rank == 0:
        off_all, mt_all, _ = ctx.match_pairs(pairs, m.MatchOptions())
        same_counts = np.array_equal(np.diff(off_all), counts.cpu().numpy())
        mine_from_all = np.concatenate([mt_all[off_all[p]:off_all[p + 1]] for p in range(0, len(pairs), world)] or [np.zeros((0, 2), np.int32)])
        same_lists = np.array_equal(mine_from_all, mt)
        print("matching sharded vs solo: counts", same_counts, "lists", same_lists, "total", int(counts.sum()), flush=True)
        ok &= same_counts and same_lists
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI-GPU", "OK" if int(flag) == 1 else "FAILED", flush=True)
    ctx.close()
    dist.destroy_process_group()
    return 0 if int(flag) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
