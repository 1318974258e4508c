"""GPU parity (through the C-ABI) of the key-frame flavoured searches against oracle/match_kf_oracle.cpp: bit-exact index and distance arrays.
Relocalisation / loop-closing / fusion projection searches, KF-KF SearchByBoW (a15) and SearchForTriangulation (f4)."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import ORBmatcher, OrbError, capi
import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n_kp,th", [(1, (1000, 900), 4.0), (2, (300, 0), 10.0), (3, (2000, 1500), 3.0)])
@pytest.mark.parametrize("quirk", [True, False])
def test_project_best(seed, n_kp, th, quirk):
    frame, view, pts, _ = synth.kf_projection_scene(seed, n_kp=n_kp)
    m = ORBmatcher()
    for variant in (capi.KF_SEARCH, capi.KF_FUSE, capi.KF_FUSE_SIM3):
        bk, bd = m.ProjectBest(frame, view, pts, th, variant, kf_index_quirk=quirk)
        rk, rd = O.project_best(frame, view, pts, th, variant, kf_quirk=quirk)
        assert np.array_equal(bk, rk) and np.array_equal(bd, rd), (variant, (bk != rk).sum())
        assert (bk >= 0).sum() > 50


@pytest.mark.parametrize("seed", [4, 5])
@pytest.mark.parametrize("cam", [0, 1])
@pytest.mark.parametrize("ori", [True, False])
def test_reloc_search(seed, cam, ori):
    frame, view, pts, blocked = synth.kf_projection_scene(seed)
    m = ORBmatcher(0.9, ori)
    for th, orb_dist in ((10.0, 100), (3.0, 64)):            # the two calls of Tracking::Relocalization (src/Tracking.cc:934, 948)
        n, out = m.SearchByProjectionReloc(frame, view, cam, pts, th, orb_dist, blocked)
        rn, rout = O.search_by_projection_reloc(frame, view, cam, pts, th, orb_dist, blocked, check_ori=ori)
        assert n == rn and np.array_equal(out, rout)
        assert n > 30


@pytest.mark.parametrize("cam", [0, 1])
@pytest.mark.parametrize("quirk", [True, False])
def test_sim3_search(cam, quirk):
    frame, view, pts, _ = synth.kf_projection_scene(6)
    matched = (np.random.default_rng(9).random(int(frame["n_kp"][cam])) < 0.2).astype(np.uint8)
    m = ORBmatcher(0.75, True)
    n, loc = m.SearchByProjectionSim3(frame, view, cam, pts, 10, matched, kf_index_quirk=quirk)
    rn, rloc = O.search_by_projection_sim3(frame, view, cam, pts, 10, matched, kf_quirk=quirk)
    assert n == rn and np.array_equal(loc, rloc)
    if cam == 0 or not quirk:
        assert n > 30


def test_projection_edge_cases():
    frame, view, pts, blocked = synth.kf_projection_scene(7, n_kp=(50, 40), n_stray=10)
    m = ORBmatcher()
    empty = {k: v[:0] for k, v in pts.items()}
    bk, bd = m.ProjectBest(frame, view, empty, 4.0, capi.KF_FUSE)
    assert bk.shape == (2, 0)
    n, out = m.SearchByProjectionReloc(frame, view, 0, empty, 10.0, 100, blocked)
    assert n == 0 and (out == -1).all()
    none = dict(pts, valid=np.zeros_like(pts["valid"]))
    bk, bd = m.ProjectBest(frame, view, none, 4.0, capi.KF_SEARCH)
    assert (bk == -1).all() and (bd == 256).all()
    nokp = dict(frame, n_kp=np.zeros(2, np.int32), kps_un=frame["kps_un"][:0], desc=frame["desc"][:0])
    bk, bd = m.ProjectBest(nokp, view, pts, 4.0, capi.KF_SEARCH)
    assert (bk == -1).all()
    with pytest.raises(OrbError):
        m.ProjectBest(frame, view, pts, 4.0, 7)
    with pytest.raises(OrbError):
        m.SearchByProjectionReloc(frame, view, 2, pts, 10.0, 100, blocked)
    with pytest.raises(OrbError):
        m.ProjectBest(frame, view, {k: v for k, v in pts.items() if k != "normal"}, 4.0, capi.KF_FUSE)     # the 60 degree test needs normals


@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("ori", [True, False])
def test_bow_kf(seed, ori):
    F, KF, valid = synth.bow_scene(seed)
    v2 = (np.random.default_rng(seed).random(int(F["n_kp"].sum())) < 0.7).astype(np.uint8)
    m = ORBmatcher(0.75, ori)
    for c1, c2 in ((0, 0), (1, 1), (0, 1), (1, 0)):           # same-camera and cross-camera calls (src/LoopClosing.cc:300, src/Tracking.cc:822)
        n, m12 = m.SearchByBoWKF(KF, c1, F, c2, valid, v2)
        rn, r12 = O.search_by_bow_kf(KF, c1, F, c2, valid, v2, 0.75, check_ori=ori)
        assert n == rn and np.array_equal(m12, r12)
    assert n >= 0


@pytest.mark.parametrize("seed,cam", [(3, 0), (4, 1), (8, 0)])
@pytest.mark.parametrize("ori", [True, False])
def test_search_for_triangulation(seed, cam, ori):
    s = synth.triangulation_scene(seed, cam=cam)
    args = (s["K1"], s["K2"], s["cam"], s["kps1"], s["kps2"], s["has_mp1"], s["has_mp2"], s["F12"], s["C1sw"], s["R2sw"], s["t2sw"], s["K2cam"], s["scale_factors"])
    m = ORBmatcher(0.6, ori)
    n, m12 = m.SearchForTriangulation(*args)
    rn, r12 = O.search_for_triangulation(*args, check_ori=ori)
    assert n == rn and np.array_equal(m12, r12)
    assert n > 60


def test_bow_frame_kf_still_matches_after_refactor():
    F, KF, valid = synth.bow_scene(11)
    m = ORBmatcher(0.7, True)
    for scaled in (True, False):
        n, out = m.SearchByBoW(F, KF, valid, bMapScaled=scaled)
        rn, rout = O.search_by_bow(F, KF, valid, 0.7, True, scaled)
        assert n == rn and np.array_equal(out, rout)


def test_project_best_per_camera_sequencing():
    """the camera loop of SearchByProjection(pKF, vpMapPoints, sFound, ...) (src/ORBmatcher.cc:703-797): points matched in camera 0 get their
    descriptor / normal / depth range refreshed before camera 1 projects them.  The per-camera entry point reproduces that: camera 1 searched
    with the refreshed fields equals the oracle run on the refreshed fields, and differs from the one-shot (snapshot) call."""
    frame, view, pts, _ = synth.kf_projection_scene(11)
    m = ORBmatcher()
    bk0, bd0 = m.ProjectBestCam(frame, view, pts, 4.0, capi.KF_SEARCH, cam=0, kf_index_quirk=False)
    all_k, all_d = m.ProjectBest(frame, view, pts, 4.0, capi.KF_SEARCH, kf_index_quirk=False)
    assert np.array_equal(bk0, all_k[0]) and np.array_equal(bd0, all_d[0])
    # "ComputeDistinctiveDescriptors / UpdateNormalAndDepth" between the cameras: the points that camera 1 is going to match get a new
    # descriptor (another key point's) and a different depth range, as if camera 0 had just matched them
    hit = np.flatnonzero((all_k[1] >= 0) & (all_d[1] <= 100))
    assert len(hit) > 20
    refreshed = {k: np.array(v, copy=True) for k, v in pts.items()}
    refreshed["desc"][hit] = frame["desc"][(all_k[1][hit] + 7) % len(frame["desc"])]
    refreshed["max_dist"][hit] *= 1.1
    refreshed["min_dist"][hit] *= 0.9
    bk1, bd1 = m.ProjectBestCam(frame, view, refreshed, 4.0, capi.KF_SEARCH, cam=1, kf_index_quirk=False)
    rk, rd = O.project_best(frame, view, refreshed, 4.0, capi.KF_SEARCH, kf_quirk=False)
    assert np.array_equal(bk1, rk[1]) and np.array_equal(bd1, rd[1])
    assert not (np.array_equal(bk1, all_k[1]) and np.array_equal(bd1, all_d[1])), "the refresh should change camera 1's result for some point"
    with pytest.raises(OrbError):
        m.ProjectBestCam(frame, view, pts, 4.0, capi.KF_SEARCH, cam=2)
