"""GPU parity of the large / distributed GlobalBundleAdjustemnt path (orbba_dist_*, world = 1 here; tools/gba_dist_run.py covers
world > 1 on a multi-GPU box) against the FP64 CPU oracle.  Bar: poses within 1e-5 relative."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import DistributedOptimizer, Optimizer, shard_problem, synth

pytestmark = pytest.mark.gpu


def _pose_rel(a, b):
    a, b = a.reshape(-1, 12), b.reshape(-1, 12)
    return (np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max()


@pytest.mark.parametrize("kw,its,robust", [(dict(seed=1, n_kf=40, n_points=2500), 10, True), (dict(seed=2, n_kf=25, n_points=900, outlier_frac=0.0), 6, False)])
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
    """a map an order of magnitude beyond the LocalBA window: 300 key frames, 30k points (dense 1794 x 1794 reduced system)"""
    p = synth.gba_problem(3, n_kf=300, n_points=30000)
    opt = DistributedOptimizer()
    poses, points, st = opt.GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=8)
    assert st["iterations"] == 8 and st["final_chi2"] < 0.5 * st["initial_chi2"]
    assert np.isfinite(poses).all() and np.isfinite(points).all()
    assert _pose_rel(poses[:1], p["poses"][:1]) < 1e-7          # the fixed pose stays put
