"""GPU parity of the guided searches (through the C-ABI) against the CPU oracle: index-exact assignments and counts.
Reference: ORBmatcher::SearchByProjection (x2), SearchByBoW, Frame::isInFrustum (src/ORBmatcher.cc, src/Frame.cc)."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import ORBmatcher
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher():
    return ORBmatcher(nnratio=0.8, checkOri=True)


@pytest.mark.parametrize("seed,n_kp,n_mp,th", [(1, (1000, 1000), 2500, 3.0), (2, (2000, 1800), 4000, 1.0), (3, (40, 0), 60, 5.0),
                                                (4, (1000, 1000), 0, 3.0), (5, (0, 0), 50, 3.0)])
def test_search_by_projection(matcher, seed, n_kp, n_mp, th):
    frame = synth.search_frame(seed, n_kp=n_kp)
    mps = synth.local_map_points(seed + 100, frame, n_mp)
    total = int(sum(n_kp))
    blocked = (np.random.default_rng(seed).random(total) < 0.15).astype(np.uint8)
    n, out = matcher.SearchByProjection(frame, mps, th=th, blocked=blocked)
    rn, rout = O.search_by_projection(frame, mps, th=th, nnratio=0.8, blocked=blocked)
    assert n == rn and np.array_equal(out, rout)
    if n_mp >= 2000:
        assert n > 500


def test_search_by_projection_unclustered_and_overwrite(matcher):
    frame = synth.search_frame(9, n_kp=(1200, 900), clustered=False)
    mps = synth.local_map_points(10, frame, 3000, frac_obs0=0.5)      # half of the map points do not block their keypoint
    n, out = matcher.SearchByProjection(frame, mps)
    rn, rout = O.search_by_projection(frame, mps, nnratio=0.8)
    assert n == rn and np.array_equal(out, rout)
    assert n > (out >= 0).sum()                                        # some keypoints were assigned more than once


@pytest.mark.parametrize("seed,scaled,ori", [(1, True, True), (2, False, True), (3, True, False)])
def test_search_by_projection_last(seed, scaled, ori):
    m = ORBmatcher(nnratio=0.9, checkOri=ori)
    cur, Rsw, tsw, K, last = synth.motion_model_scene(seed)
    blocked = (np.random.default_rng(seed).random(int(cur["n_kp"].sum())) < 0.05).astype(np.uint8)
    n, out, per_cam = m.SearchByProjectionLast(cur, Rsw, tsw, K, last, th=15.0, bMapScaled=scaled, blocked=blocked)
    rn, rout, rper = O.search_by_projection_last(cur, Rsw, tsw, K, last, th=15.0, check_ori=ori, map_scaled=scaled, blocked=blocked)
    assert n == rn and np.array_equal(per_cam, rper) and np.array_equal(out, rout)
    assert n > 200


def test_search_by_projection_last_few_matches_quirk():
    m = ORBmatcher(checkOri=True)
    cur, Rsw, tsw, K, last = synth.motion_model_scene(4)
    idx0 = np.flatnonzero(last["cam"] == 0)
    last["valid"] = last["valid"].copy()
    last["valid"][idx0[12:]] = 0
    n, out, per_cam = m.SearchByProjectionLast(cur, Rsw, tsw, K, last, th=7.0)
    rn, rout, rper = O.search_by_projection_last(cur, Rsw, tsw, K, last, th=7.0)
    assert n == rn <= 20 and np.array_equal(per_cam, rper) and np.array_equal(out, rout)
    assert (out[cur["n_kp"][0]:] == -1).all()


@pytest.mark.parametrize("seed,n_kp,nodes,ori,scaled", [(1, (1000, 900), 90, True, True), (2, (2000, 2000), 100, False, True),
                                                        (3, (300, 50), 6, True, False), (4, (500, 0), 20, True, True)])
def test_search_by_bow(seed, n_kp, nodes, ori, scaled):
    m = ORBmatcher(nnratio=0.7, checkOri=ori)
    F, KF, valid = synth.bow_scene(seed, n_kp=n_kp, n_nodes=nodes)
    n, out = m.SearchByBoW(F, KF, valid, bMapScaled=scaled)
    rn, rout = O.search_by_bow(F, KF, valid, nnratio=0.7, check_ori=ori, map_scaled=scaled)
    assert n == rn and np.array_equal(out, rout)
    assert n > 10


def test_is_in_frustum(matcher):
    frame, pos, normal, mx, mn = synth.frustum_scene(1, n=20000)
    out, uvc = matcher.isInFrustum(frame, pos, normal, mx, mn)
    rout, ruvc = O.is_in_frustum(frame, pos, normal, mx, mn)
    assert np.array_equal(out, rout)
    assert uvc.tobytes() == ruvc.tobytes()          # u, v, viewCos bit-exact (FP32 without contraction on both sides)
    out0, _ = matcher.isInFrustum(frame, pos, normal, mx, mn, bForAllCam=False)
    assert np.array_equal(out0, O.is_in_frustum(frame, pos, normal, mx, mn, for_all=False)[0])


def test_frustum_feeds_search(matcher):
    """isInFrustum -> orbm_mp_t -> SearchByProjection, the TrackLocalMap chain (src/Tracking.cc SearchLocalPoints)."""
    frame = synth.search_frame(21, n_kp=(1000, 1000))
    fr, pos, normal, mx, mn = synth.frustum_scene(22, n=3000)
    out, uvc = matcher.isInFrustum(fr, pos, normal, mx, mn)
    mps = np.zeros(len(out), synth.MP_DTYPE)
    mps["valid"] = out[:, 0]; mps["cam"] = np.maximum(out[:, 1], 0); mps["level"] = out[:, 2]
    mps["u"] = uvc[:, 0]; mps["v"] = uvc[:, 1]; mps["view_cos"] = uvc[:, 2]; mps["obs_positive"] = 1
    mps["desc"] = synth.random_descriptors(5, len(out))
    n, o = matcher.SearchByProjection(frame, mps, th=5.0)
    rn, ro = O.search_by_projection(frame, mps, th=5.0, nnratio=0.8)
    assert n == rn and np.array_equal(o, ro)
