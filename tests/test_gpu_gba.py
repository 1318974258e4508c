"""GPU parity of the large / distributed GlobalBundleAdjustemnt path (orbba_dist_*) against the FP64 CPU oracle; world > 1 is
covered by test_dist_world_n_equals_world1 (skipped on a box with one GPU).  Bar: poses within 1e-5 relative."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import DistributedOptimizer, Optimizer, shard_problem
import synth

pytestmark = pytest.mark.gpu


def _pose_rel(a, b):
    a, b = a.reshape(-1, 12), b.reshape(-1, 12)
    return (np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max()


@pytest.mark.parametrize("kw,its,robust", [(dict(seed=1, n_kf=40, n_points=2500), 10, True), (dict(seed=2, n_kf=25, n_points=900, outlier_frac=0.0), 6, False),
                                           (dict(seed=3, n_kf=300, n_points=12000), 5, True)],
                         ids=["40kf_huber", "25kf_plain", "300kf_1794x1794_reduced_system"])
def test_dist_world1_vs_oracle(kw, its, robust):
    p = synth.gba_problem(**kw)
    opt = DistributedOptimizer()
    poses, points, st = opt.GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=its, bRobust=robust)
    rc, rposes, rpoints, rst = O.global_ba(p, iterations=its, huber_delta=O.HUBER_2D if robust else 0.0)
    assert _pose_rel(poses, rposes) <= 1e-5, _pose_rel(poses, rposes)
    assert np.abs(points - rpoints).max() <= 1e-5 * max(1.0, np.abs(rpoints).max())
    assert st["iterations"] == rst["iterations"] and st["trials"] == rst["trials"]
    assert np.isclose(st["initial_chi2"], rst["initial_chi2"], rtol=1e-9) and np.isclose(st["final_chi2"], rst["final_chi2"], rtol=1e-6)


def test_dist_matches_batched_path():
    """the same problem through the batched LocalBA machinery (orbba_global) and through the large-map path"""
    p = synth.ba_problem(4, n_kf=12, n_points=500, outlier_frac=0.02)
    a, _, sa = Optimizer().GlobalBundleAdjustemnt(p, nIterations=8, bRobust=True)
    b, _, sb = DistributedOptimizer().GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=8, bRobust=True)
    assert _pose_rel(a, b) <= 1e-7 and sa["trials"] == sb["trials"]


def test_dist_large_converges():
    """a map an order of magnitude beyond the LocalBA window: 300 key frames, 30k points (1794 x 1794 reduced system, block skyline)"""
    p = synth.gba_problem(3, n_kf=300, n_points=30000)
    opt = DistributedOptimizer()
    poses, points, st = opt.GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=8)
    assert st["iterations"] == 8 and st["final_chi2"] < 0.5 * st["initial_chi2"]
    assert np.isfinite(poses).all() and np.isfinite(points).all()
    assert _pose_rel(poses[:1], p["poses"][:1]) < 1e-7          # the fixed pose stays put


def _loop_closure_problem(seed, n_kf, n_points, n_loop):
    """a band plus `n_loop` landmarks that the first and the last key frames both observe: long rows in the skyline"""
    p = synth.gba_problem(seed, n_kf=n_kf, n_points=n_points)
    rng = np.random.default_rng(seed)
    src = np.flatnonzero(p["edge_pose"] >= n_kf - 3)[:n_loop]       # edges of the last key frames ...
    extra = {k: p[k][src].copy() for k in ("edge_pose", "edge_point", "edge_cam", "edge_obs", "edge_inv_sigma2")}
    extra["edge_pose"] = rng.integers(1, 4, len(src)).astype(np.int32)   # ... observed again (with a large residual: Huber territory) from key frames 1..3
    out = dict(p)
    for k in extra:
        out[k] = np.ascontiguousarray(np.concatenate([p[k], extra[k]]))
    keep = np.ones(len(out["edge_pose"]), bool)                     # one observation per (landmark, key frame)
    seen = set()
    for e, (a, b) in enumerate(zip(out["edge_point"], out["edge_pose"])):
        if (a, b) in seen:
            keep[e] = False
        seen.add((a, b))
    for k in extra:
        out[k] = np.ascontiguousarray(out[k][keep])
    return out


def test_dist_loop_closure_envelope_vs_oracle():
    p = _loop_closure_problem(5, 60, 3000, 40)
    opt = DistributedOptimizer()
    poses, points, st = opt.GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=6, bRobust=True)
    rc, rposes, rpoints, rst = O.global_ba(p, iterations=6)
    assert _pose_rel(poses, rposes) <= 1e-5, _pose_rel(poses, rposes)
    assert st["iterations"] == rst["iterations"] and st["trials"] == rst["trials"]
    tm = opt.timing()
    assert tm["skyline_blocks"] > 59 * 8          # the long rows are in the envelope


@pytest.mark.parametrize("n_kf,segments", [(120, 2), (120, 5), (300, None), (333, 9)])
def test_dist_substructured_solve_equals_band(n_kf, segments, monkeypatch):
    """segments + separators (one CTA per segment, reduced band for the separators) against the single-CTA band factorisation of the
    same system: same LM trajectory, poses to rounding; the substructured path is itself bit-reproducible"""
    p = synth.gba_problem(11, n_kf=n_kf, n_points=40 * n_kf)
    sh = shard_problem(p, 0, 1)
    monkeypatch.setenv("ORBGBA_SEGMENTS", "0")
    ref = DistributedOptimizer()
    a = ref.GlobalBundleAdjustemnt(sh, nIterations=5)
    assert ref.timing()["segments"] == 0
    if segments is None: monkeypatch.delenv("ORBGBA_SEGMENTS")
    else: monkeypatch.setenv("ORBGBA_SEGMENTS", str(segments))
    opt = DistributedOptimizer()
    b = opt.GlobalBundleAdjustemnt(sh, nIterations=5)
    c = DistributedOptimizer().GlobalBundleAdjustemnt(sh, nIterations=5)
    assert opt.timing()["segments"] >= 2
    assert a[2]["trials"] == b[2]["trials"] and a[2]["iterations"] == b[2]["iterations"]
    assert _pose_rel(a[0], b[0]) <= 1e-9 and np.isclose(a[2]["final_chi2"], b[2]["final_chi2"], rtol=1e-9)
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1])


def test_dist_bit_reproducible():
    p = synth.gba_problem(6, n_kf=50, n_points=3000)
    sh = shard_problem(p, 0, 1)
    a = DistributedOptimizer().GlobalBundleAdjustemnt(sh, nIterations=4)
    b = DistributedOptimizer().GlobalBundleAdjustemnt(sh, nIterations=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])       # every sum has one owner and a fixed order


def test_dist_world_n_equals_world1():
    """landmark-partitioned over every GPU of the box (one process per GPU, NCCL all-reduce of the skyline): same poses as one GPU"""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(root, "tools", "gba_dist_run.py"), "200", "12000", "4", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == n and res["check"]["trials_equal"] and res["check"]["pose_rel_vs_1gpu"] <= 1e-9, res
