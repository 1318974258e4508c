"""GPU parity of Optimizer::PoseOptimization (orbba_pose_optimization) against the FP64 oracle: pose within 1e-5 relative,
identical outlier flags, inlier count and LM iteration / trial counts."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import Optimizer
import synth

pytestmark = pytest.mark.gpu


def _check(got, ref):
    pose, out, inl, cnt = got
    rpose, rout, rinl, rcnt = ref
    rel = np.linalg.norm(pose - rpose) / np.linalg.norm(rpose)
    # the north-star bar is 1e-5; observed 1e-16.  With fewer than 3 correspondences nothing is optimised and the oracle hands the CV_32F input
    # back untouched, while the library returns it after the SE3Quat round trip (a float-rounded rotation is orthonormal to 1e-8 only)
    assert rel <= (1e-9 if len(out) >= 3 else 1e-7), rel
    assert np.array_equal(out, rout) and inl == rinl
    # LM iteration / trial counts are the oracle's.  The one tolerated deviation: after a round has converged to the last bit, rho =
    # (chi - chi') / scale is pure rounding noise and its sign decides between "terminate (rho == 0)" and "one more iteration", so the
    # parallel sums of the GPU may take up to two no-op iterations more or fewer than the sequential sums of the oracle -- only with
    # the estimate agreeing to 1e-12 (seed 3 of the list below: 23 vs 21 iterations, equal trials, pose equal to 2e-16).
    if tuple(cnt) != tuple(rcnt):
        assert rel <= 1e-12 and abs(cnt[0] - rcnt[0]) <= 2 and abs(cnt[1] - rcnt[1]) <= 2, (cnt, rcnt, rel)


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
