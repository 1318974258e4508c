"""CPU checks of the key-frame flavoured search oracle (oracle/match_kf_oracle.cpp): brute-force numpy restatements of the
decision rules on small scenes, and properties (claims are exclusive, thresholds hold, camera-0 results do not depend on the
KeyFrame::GetFeaturesInArea index quirk)."""
import numpy as np

import oracle_lib as O
import synth


def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def test_project_best_variants_are_consistent():
    frame, view, pts, _ = synth.kf_projection_scene(1, n_kp=(300, 250), n_stray=60)
    first = np.concatenate([[0], np.cumsum(frame["n_kp"])])
    for variant in (0, 1, 2):
        bk, bd = O.project_best(frame, view, pts, 4.0, variant, kf_quirk=False)
        assert bk.shape == (2, len(pts["valid"]))
        hit = bk >= 0
        assert hit.sum() > 50
        assert (bd[~hit] == 256).all() and (bd[hit] < 256).all()
        for c in range(2):
            assert ((bk[c][hit[c]] >= first[c]) & (bk[c][hit[c]] < first[c + 1])).all()          # the key point belongs to the searched camera
            for i in np.flatnonzero(hit[c])[:40]:
                assert bd[c, i] == _ham(pts["desc"][i], frame["desc"][bk[c, i]])
        assert not hit[:, pts["valid"] == 0].any()
    # Fuse adds the chi-square pixel gate to the Sim3 variant's gates: every Fuse hit is a Sim3 candidate at no larger distance
    b1, d1 = O.project_best(frame, view, pts, 4.0, 1, kf_quirk=False)
    b2, d2 = O.project_best(frame, view, pts, 4.0, 2, kf_quirk=False)
    assert ((b1 >= 0) <= (b2 >= 0)).all() and (d2[b1 >= 0] <= d1[b1 >= 0]).all()
    # the index quirk of KeyFrame::GetFeaturesInArea only concerns cameras > 0
    q0 = O.project_best(frame, view, pts, 4.0, 0, kf_quirk=True)
    q1 = O.project_best(frame, view, pts, 4.0, 0, kf_quirk=False)
    assert np.array_equal(q0[0][0], q1[0][0]) and np.array_equal(q0[1][0], q1[1][0])
    assert not np.array_equal(q0[0][1], q1[0][1])


def test_reloc_and_sim3_claims():
    frame, view, pts, blocked = synth.kf_projection_scene(2, n_kp=(400, 300), n_stray=80)
    for cam in (0, 1):
        n, out = O.search_by_projection_reloc(frame, view, cam, pts, 10.0, 100, blocked, check_ori=False)
        m = out >= 0
        assert n == m.sum() and n > 30
        assert not (m & (blocked != 0)).any()                                # key points that already hold a map point are skipped
        assert len(np.unique(out[m])) == n                                   # one key point per map point
        assert all(_ham(pts["desc"][out[g]], frame["desc"][g]) <= 100 for g in np.flatnonzero(m))
        n2, out2 = O.search_by_projection_reloc(frame, view, cam, pts, 10.0, 100, blocked, check_ori=True)
        assert n2 <= n and ((out2 >= 0) <= m).all()                          # the rotation check only removes
        nl = int(frame["n_kp"][cam])
        matched = (np.random.default_rng(3).random(nl) < 0.2).astype(np.uint8)
        n3, loc = O.search_by_projection_sim3(frame, view, cam, pts, 10, matched, kf_quirk=False)
        m3 = loc >= 0
        assert n3 == m3.sum() and not (m3 & (matched != 0)).any()
        first = int(np.sum(frame["n_kp"][:cam]))
        assert all(_ham(pts["desc"][loc[l]], frame["desc"][first + l]) <= 50 for l in np.flatnonzero(m3))
        assert n3 > 10


def test_bow_kf_and_triangulation():
    F, KF, valid = synth.bow_scene(5, n_kp=(400, 350), n_nodes=40)
    v2 = (np.random.default_rng(1).random(int(F["n_kp"].sum())) < 0.7).astype(np.uint8)
    for c1, c2 in ((0, 0), (1, 1), (0, 1)):
        n, m12 = O.search_by_bow_kf(KF, c1, F, c2, valid, v2, 0.75, check_ori=False)
        hit = m12 >= 0
        assert n == hit.sum() and len(np.unique(m12[hit])) == n
        firstK = int(np.sum(KF["n_kp"][:c1])); firstF = int(np.sum(F["n_kp"][:c2]))
        assert ((m12[hit] >= firstF) & (m12[hit] < firstF + F["n_kp"][c2])).all()
        assert v2[m12[hit]].all() and valid[firstK + np.flatnonzero(hit)].all()
        assert all(_ham(KF["desc"][firstK + l], F["desc"][m12[l]]) < 50 for l in np.flatnonzero(hit))
        if c1 == c2:
            assert n > 40
    s = synth.triangulation_scene(3, n=400)
    args = (s["K1"], s["K2"], s["cam"], s["kps1"], s["kps2"], s["has_mp1"], s["has_mp2"], s["F12"], s["C1sw"], s["R2sw"], s["t2sw"], s["K2cam"], s["scale_factors"])
    n, m12 = O.search_for_triangulation(*args, check_ori=False)
    hit = m12 >= 0
    assert n == hit.sum() and n > 40 and len(np.unique(m12[hit])) == n
    assert not s["has_mp1"][np.flatnonzero(hit)].any() and not s["has_mp2"][m12[hit]].any()       # camera 0: local == global
    # the epipolar gate: accepted pairs satisfy x1^T F12 x2 ~ 0 within 3.84 sigma^2
    F12 = s["F12"].reshape(3, 3).astype(np.float64)
    for l in np.flatnonzero(hit)[:60]:
        k1, k2 = s["kps1"][l], s["kps2"][m12[l]]
        line = np.array([k1["x"], k1["y"], 1.0]) @ F12
        dsq = (line @ np.array([k2["x"], k2["y"], 1.0])) ** 2 / (line[0] ** 2 + line[1] ** 2)
        assert dsq < 3.84 * float(s["scale_factors"][k2["octave"]]) ** 2 * 1.001
    n2, m2 = O.search_for_triangulation(*args, check_ori=True)
    assert n2 <= n and n2 > 20
