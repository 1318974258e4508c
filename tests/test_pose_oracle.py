"""CPU checks of the PoseOptimization oracle (Optimizer::PoseOptimization, src/Optimizer.cc:250-405)."""
import numpy as np

import oracle_lib as O
import synth


def test_pose_optimization_recovers_pose_and_outliers():
    f = synth.pose_opt_frame(1, n_obs=500)
    pose, out, inl, (its, trials) = O.pose_optimization(f)
    e0 = np.linalg.norm(f["pose"] - f["gt_pose"]); e1 = np.linalg.norm(pose - f["gt_pose"])
    assert e1 < 0.1 * e0 and e1 < 0.02
    assert inl == (~out).sum() and its >= 4 and trials >= its
    assert (out & f["planted"]).sum() >= 0.95 * f["planted"].sum()      # the gross outliers are found
    assert (out & ~f["planted"]).sum() <= 0.12 * (~f["planted"]).sum()  # chi2 > 5.991 rejects ~5 % of clean edges by design


def test_pose_optimization_small_inputs():
    f = synth.pose_opt_frame(2, n_obs=2)
    pose, out, inl, _ = O.pose_optimization(f)
    assert inl == 0 and np.array_equal(pose, f["pose"]) and not out.any()          # fewer than 3 correspondences: nothing is touched
    g = synth.pose_opt_frame(3, n_obs=8, outlier_frac=0.0)
    pose, out, inl, (its, trials) = O.pose_optimization(g)
    assert its <= 10 and inl <= 8                                                   # fewer than 10 edges: a single round
