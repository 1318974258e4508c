"""GPU parity of Optimizer::PoseOptimization (orbba_pose_optimization) against the FP64 oracle: pose within 1e-5 relative,
identical outlier flags, inlier count and LM iteration / trial counts."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import Optimizer, synth

pytestmark = pytest.mark.gpu


def _check(got, ref):
    pose, out, inl, cnt = got
    rpose, rout, rinl, rcnt = ref
    assert np.linalg.norm(pose - rpose) / np.linalg.norm(rpose) <= 1e-5
    assert np.array_equal(out, rout) and inl == rinl
    # LM iteration / trial counts: once a round has converged, rho = (chi - chi') / scale is rounding noise (its sign decides between
    # "accept", "rho == 0 -> terminate" and "retry"), so the counts may differ by a few steps there; the estimates do not
    assert abs(cnt[0] - rcnt[0]) <= 4 and abs(cnt[1] - rcnt[1]) <= 12, (cnt, rcnt)


@pytest.mark.parametrize("kw", [dict(seed=1, n_obs=600), dict(seed=2, n_obs=2000, outlier_frac=0.3), dict(seed=3, n_obs=40, outlier_frac=0.05),
                                dict(seed=4, n_obs=9), dict(seed=5, n_obs=2), dict(seed=6, n_obs=0), dict(seed=7, n_obs=300, pose_noise=(0.3, 8.0))])
def test_pose_optimization_vs_oracle(kw):
    f = synth.pose_opt_frame(**kw)
    _check(Optimizer().PoseOptimization(f), O.pose_optimization(f))


def test_pose_optimization_batch():
    frames = [synth.pose_opt_frame(20 + i, n_obs=200 + 150 * i) for i in range(9)]
    opt = Optimizer()
    res = opt.PoseOptimization(frames)
    for f, r in zip(frames, res):
        _check(r, O.pose_optimization(f))
    again = opt.PoseOptimization(frames)
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(res, again))      # bit-reproducible
