"""CPU checks of the guided-search oracle (oracle/match_oracle.cpp) against independent plain-Python restatements of the
reference loops on small inputs, and of its structural properties on larger ones."""
import numpy as np

import oracle_lib as O
import synth


def _hamming(a, b):
    return int(np.unpackbits(a ^ b).sum())


def _py_features_in_area(frame, c, x, y, r, lo, hi):
    """brute force over all keypoints of camera c, then ordered like mvGrids[c][ix][iy] (column-major cells, insertion order)"""
    first = np.concatenate([[0], np.cumsum(frame["n_kp"])])
    b = frame["bounds"][c]
    invW = np.float32(64) / np.float32(b[1] - b[0]); invH = np.float32(48) / np.float32(b[3] - b[2])
    x0 = max(0, int(np.floor(np.float32(np.float32(x - b[0]) - r) * invW))); x1 = min(63, int(np.ceil(np.float32(np.float32(x - b[0]) + r) * invW)))
    y0 = max(0, int(np.floor(np.float32(np.float32(y - b[2]) - r) * invH))); y1 = min(47, int(np.ceil(np.float32(np.float32(y - b[2]) + r) * invH)))
    res = []
    for i in range(frame["n_kp"][c]):
        kp = frame["kps_un"][first[c] + i]
        px = int(np.rint(np.float32(kp["x"] - b[0]) * invW)); py = int(np.rint(np.float32(kp["y"] - b[2]) * invH))
        if not (0 <= px < 64 and 0 <= py < 48) or not (x0 <= px <= x1 and y0 <= py <= y1):
            continue
        if kp["octave"] < lo or kp["octave"] > hi:
            continue
        if abs(np.float32(kp["x"] - x)) < r and abs(np.float32(kp["y"] - y)) < r:
            res.append((px, py, i))
    return [i for _, _, i in sorted(res)]


def test_search_by_projection_matches_python_restatement():
    frame = synth.search_frame(3, n_kp=(150, 120))
    mps = synth.local_map_points(4, frame, 200)
    n, out = O.search_by_projection(frame, mps, th=3.0, nnratio=0.8)
    first = np.concatenate([[0], np.cumsum(frame["n_kp"])])
    blocked = np.zeros(first[-1], bool)
    ref = np.full(first[-1], -1)
    cnt = 0
    sf = frame["scale_factors"]
    for i, mp in enumerate(mps):
        if not mp["valid"]:
            continue
        r = np.float32(2.5 if mp["view_cos"] > 0.998 else 4.0) * np.float32(3.0)
        r = np.float32(r * sf[mp["level"]])
        cand = _py_features_in_area(frame, mp["cam"], mp["u"], mp["v"], r, mp["level"] - 1, mp["level"] + 1)
        best = (256, -1, -1); second = (256, -1)
        for l in cand:
            g = first[mp["cam"]] + l
            if blocked[g]:
                continue
            d = _hamming(mp["desc"], frame["desc"][g])
            if d < best[0]:
                second = (best[0], best[1]); best = (d, frame["kps_un"][g]["octave"], g)
            elif d < second[0]:
                second = (d, frame["kps_un"][g]["octave"])
        if best[0] <= 100:
            if best[1] == second[1] and best[0] > np.float32(0.8) * np.float32(second[0]):
                continue
            ref[best[2]] = i; blocked[best[2]] = bool(mp["obs_positive"]); cnt += 1
    assert n == cnt and np.array_equal(out, ref)
    assert cnt > 50


def test_search_by_projection_edge_cases():
    frame = synth.search_frame(5, n_kp=(300, 0))
    mps = synth.local_map_points(6, frame, 100)
    n0, out0 = O.search_by_projection(frame, mps[:0])
    assert n0 == 0 and (out0 == -1).all()
    n1, out1 = O.search_by_projection(frame, mps, blocked=np.ones(300, np.uint8))
    assert n1 == 0 and (out1 == -1).all()
    n2, out2 = O.search_by_projection(frame, mps, th=1.0)
    n3, out3 = O.search_by_projection(frame, mps, th=3.0)
    assert n3 >= n2 > 0
    assert len(np.unique(out3[out3 >= 0])) == (out3 >= 0).sum() or (mps["obs_positive"] == 0).any()


def test_motion_model_search_properties():
    cur, Rsw, tsw, K, last = synth.motion_model_scene(2)
    n, out, per_cam = O.search_by_projection_last(cur, Rsw, tsw, K, last, th=15.0)
    assert per_cam[0] > 20 and n == per_cam.sum() and n > 300
    assert (out >= 0).sum() <= n            # a keypoint can be counted twice (overwrite), never the opposite
    m = out >= 0
    first = np.concatenate([[0], np.cumsum(cur["n_kp"])])
    cam_of_kp = np.searchsorted(first, np.flatnonzero(m), side="right") - 1
    assert np.array_equal(cam_of_kp, last["cam"][out[m]])        # a map point stays in the camera that saw it
    n_mono, out_mono, pc = O.search_by_projection_last(cur, Rsw, tsw, K, last, th=15.0, map_scaled=False)
    assert n_mono == per_cam[0] and (out_mono[first[1]:] == -1).all()
    n_noori, _, _ = O.search_by_projection_last(cur, Rsw, tsw, K, last, th=15.0, check_ori=False)
    assert n_noori > n                                             # the rotation histogram removes the planted outliers
    # <= 20 matches in camera 0 stops the loop and REPLACES the total (src/ORBmatcher.cc:664-667)
    few = dict(last)
    few["valid"] = last["valid"].copy()
    idx0 = np.flatnonzero(last["cam"] == 0)
    few["valid"][idx0[10:]] = 0
    n_few, out_few, pc_few = O.search_by_projection_last(cur, Rsw, tsw, K, few, th=15.0)
    assert n_few == pc_few[0] <= 20 and pc_few[1] == 0 and (out_few[first[1]:] == -1).all()


def test_search_by_bow_matches_python_restatement():
    F, KF, valid = synth.bow_scene(1, n_kp=(120, 100), n_nodes=12)
    n, out = O.search_by_bow(F, KF, valid, nnratio=0.7, check_ori=False)
    firstF = np.concatenate([[0], np.cumsum(F["n_kp"])]); firstK = np.concatenate([[0], np.cumsum(KF["n_kp"])])
    ref = np.full(firstF[-1], -1)
    cnt = 0
    for c in range(2):
        fn = {int(F["node_id"][k]): k for k in range(F["node_first"][c], F["node_first"][c + 1])}
        inner = {}
        for k in range(KF["node_first"][c], KF["node_first"][c + 1]):
            nid = int(KF["node_id"][k])
            if nid not in fn:
                continue
            kk = fn[nid]
            for a in KF["idx"][KF["node_off"][k]:KF["node_off"][k + 1]]:
                g = firstK[c] + a
                if not valid[g]:
                    continue
                b1, b2, bi = 256, 256, -1
                for l in F["idx"][F["node_off"][kk]:F["node_off"][kk + 1]]:
                    if l in inner:
                        continue
                    d = _hamming(KF["desc"][g], F["desc"][firstF[c] + l])
                    if d < b1:
                        b2, b1, bi = b1, d, l
                    elif d < b2:
                        b2 = d
                if b1 <= 50 and np.float32(b1) < np.float32(0.7) * np.float32(b2):
                    inner[bi] = g; cnt += 1
        for l, g in inner.items():
            ref[firstF[c] + l] = g
    assert n == cnt and np.array_equal(out, ref) and cnt > 20
    n2, out2 = O.search_by_bow(F, KF, valid, nnratio=0.7, check_ori=True)
    assert 0 < n2 <= n and ((out2 >= 0) <= (out >= 0)).all()
    n3, out3 = O.search_by_bow(F, KF, valid, map_scaled=False)
    assert (out3[firstF[1]:] == -1).all()


def test_is_in_frustum_properties():
    frame, pos, normal, mx, mn = synth.frustum_scene(0, n=2000)
    out, uvc = O.is_in_frustum(frame, pos, normal, mx, mn)
    vis = out[:, 0] == 1
    assert 100 < vis.sum() < 1900
    assert (uvc[vis, 0] >= 0).all() and (uvc[vis, 0] <= 640).all() and (uvc[vis, 2] >= 0.5).all()
    assert set(np.unique(out[vis, 1])) <= {0, 1} and (out[vis, 2] >= 0).all() and (out[vis, 2] <= 7).all()
    out0, _ = O.is_in_frustum(frame, pos, normal, mx, mn, for_all=False)
    assert (out0[out0[:, 0] == 1, 1] == 0).all() and (out0[:, 0] <= out[:, 0]).all()
