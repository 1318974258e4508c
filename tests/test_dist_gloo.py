"""World-size-2 check of the multi-GPU host logic on CPU (gloo): the landmark partition of shard_problem() and the exchange step
of the distributed GlobalBA -- every rank builds the Schur complement of ITS landmarks, an all-reduce(sum) makes the reduced camera
system equal to the one built from the whole map.  The per-shard normal equations come from the CPU oracle (checker only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reduced_system(p, lam):
    import oracle_lib as O
    Hpp, bp, Hll, bl, Hpl, chi2 = O.ba_normal_equations(p)
    free = np.flatnonzero(p["pose_fixed"] == 0)
    slot = -np.ones(len(p["pose_fixed"]), int)
    slot[free] = np.arange(len(free))
    K = len(free)
    Hs = np.zeros((6 * K, 6 * K)); bs = np.zeros(6 * K)
    Dinv = np.linalg.inv(Hll + lam * np.eye(3))
    by_pt = {}
    for e, (l, k) in enumerate(zip(p["edge_point"], slot[p["edge_pose"]])):
        if k >= 0:
            by_pt.setdefault(int(l), []).append((e, int(k)))
    for l, lst in by_pt.items():
        for ea, ka in lst:
            Ya = Hpl[ea] @ Dinv[l]
            bs[6 * ka:6 * ka + 6] -= Ya @ bl[l]
            for ec, kc in lst:
                Hs[6 * ka:6 * ka + 6, 6 * kc:6 * kc + 6] -= Ya @ Hpl[ec].T
    return Hs, bs, Hpp, bp, chi2, K


def _worker(rank, world, port, q):
    for pth in (ROOT, os.path.join(ROOT, "tests")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    from orbslam2_dualcam_b200 import shard_problem
    import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.ba_problem(7, n_kf=6, n_points=90)
    lam = 0.37
    sh = shard_problem(p, rank, world)
    Hs, bs, Hpp, bp, chi2, K = _reduced_system(sh, lam)
    buf = torch.from_numpy(np.concatenate([Hs.ravel(), bs, Hpp.ravel(), bp.ravel(), [chi2, len(sh["points"]), len(sh["edge_pose"])]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        fHs, fbs, fHpp, fbp, fchi2, fK = _reduced_system(p, lam)
        full = np.concatenate([fHs.ravel(), fbs, fHpp.ravel(), fbp.ravel(), [fchi2, len(p["points"]), len(p["edge_pose"])]])
        got = buf.numpy()
        scale = np.abs(full).max()
        q.put((float(np.abs(got - full).max() / scale), int(K), int(fK)))
    dist.destroy_process_group()


def test_landmark_partition_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    err, K, fK = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert K == fK and err < 1e-12, err


def test_shard_problem_covers_everything_once():
    for pth in (ROOT,):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    from orbslam2_dualcam_b200 import shard_problem
    import synth
    p = synth.ba_problem(3, n_kf=5, n_points=101)
    seen_pts, seen_edges = [], []
    for r in range(4):
        sh = shard_problem(p, r, 4)
        seen_pts.append(sh["point_ids"]); seen_edges.append(sh["edge_ids"])
        assert np.array_equal(sh["points"], p["points"][sh["point_ids"]])
        assert np.array_equal(sh["point_ids"][sh["edge_point"]], p["edge_point"][sh["edge_ids"]])
        assert np.array_equal(sh["poses"], p["poses"])
    assert np.array_equal(np.sort(np.concatenate(seen_pts)), np.arange(101))
    assert np.array_equal(np.sort(np.concatenate(seen_edges)), np.arange(len(p["edge_pose"])))
