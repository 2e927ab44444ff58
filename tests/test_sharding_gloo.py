"""CPU, world_size 2 over gloo: the N>1 host logic — pair sharding covers every pair exactly once, and the
partial reduced camera systems of a point-sharded BA problem sum (all-reduce) to the full system."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monocularsfm_b200.sharding import shard_ba_problem, shard_pairs
    from oracle import ba_oracle as bo
    # ---- pairs
    pairs = np.array([(i, j) for i in range(9) for j in range(i)], np.int32)
    mine = shard_pairs(pairs, rank, world)
    flag = torch.zeros(len(pairs), dtype=torch.int64)
    for a, b in mine:
        flag[np.nonzero((pairs[:, 0] == a) & (pairs[:, 1] == b))[0][0]] += 1
    dist.all_reduce(flag)
    ok_pairs = bool((flag == 1).all())
    # ---- BA
    P = bo.make_problem(6, 80, 4, 9)
    L = shard_ba_problem(P, rank, world)
    r, J = bo.residual_jacobian_jets(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["fx"], L["fy"])
    U, gc, V, gp, W = bo.build_normal_equations(r, J, L["obs_cam"], L["obs_pt"], len(L["cams"]), len(L["pts"]), L["cam_const"])
    # undamped camera diagonal must be added AFTER the sum (it needs the summed diag U): use inv_radius = 0 here
    S, rhs, _, _ = bo.schur_reduce(U, gc, V, gp, W, L["obs_cam"], L["obs_pt"], L["cam_const"], 0.0)
    buf = torch.from_numpy(np.concatenate([S.reshape(-1), rhs, [bo.cost_of(r)]]))
    dist.all_reduce(buf)
    # ---- BA with the shared focal block (refine_focal_length): the bordered system [[S, B], [B^T, F]] is additive too
    r11, J11 = bo.residual_jacobian_jets_focal(L["cams"], L["pts"], L["obs_uv"], L["obs_cam"], L["obs_pt"], L["fx"], L["fy"])
    Rl, rhsl, _, _ = bo.reduced_system_focal(r11, J11, L["obs_cam"], L["obs_pt"], len(L["cams"]), len(L["pts"]), L["cam_const"], 0.0)
    buf_f = torch.from_numpy(np.concatenate([Rl.reshape(-1), rhsl]))
    dist.all_reduce(buf_f)
    if rank == 0:
        r, J = bo.residual_jacobian_jets(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
        U, gc, V, gp, W = bo.build_normal_equations(r, J, P["obs_cam"], P["obs_pt"], len(P["cams"]), len(P["pts"]), P["cam_const"])
        S, rhs, _, _ = bo.schur_reduce(U, gc, V, gp, W, P["obs_cam"], P["obs_pt"], P["cam_const"], 0.0)
        full = np.concatenate([S.reshape(-1), rhs, [bo.cost_of(r)]])
        r11, J11 = bo.residual_jacobian_jets_focal(P["cams"], P["pts"], P["obs_uv"], P["obs_cam"], P["obs_pt"], P["fx"], P["fy"])
        Rf, rhsf, _, _ = bo.reduced_system_focal(r11, J11, P["obs_cam"], P["obs_pt"], len(P["cams"]), len(P["pts"]), P["cam_const"], 0.0)
        full_f = np.concatenate([Rf.reshape(-1), rhsf])
        q.put((ok_pairs, float(np.abs(buf.numpy() - full).max() / np.abs(full).max()),
               float(np.abs(buf_f.numpy() - full_f).max() / np.abs(full_f).max())))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_pairs, err, err_focal = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_pairs
    assert err < 1e-12
    assert err_focal < 1e-12
