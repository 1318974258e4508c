"""Host-side mirror of the reference's ORBmatcher Hamming searches (include/ORBmatcher.h) above the C-ABI."""
import ctypes as C

import numpy as np

from . import capi
from .capi import stream_handle as _stream_handle
from .capi import check, lib, ptr


class ORBmatcher:
    TH_LOW = 50         # src/ORBmatcher.cc:58
    TH_HIGH = 100       # src/ORBmatcher.cc:57
    HISTO_LENGTH = 30   # src/ORBmatcher.cc:59

    def __init__(self, nnratio=0.6, checkOri=True, max_pairs=1, max_query=4096, max_train=4096, device=0):
        self._h = None
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        h = C.c_void_p()
        check(lib().orbm_create(C.byref(h), device, max_pairs, max_query, max_train))
        self._h = h
        self.max_pairs, self.max_query, self.max_train = max_pairs, max_query, max_train

    def close(self):
        if getattr(self, "_h", None):
            lib().orbm_destroy(self._h)
            self._h = None

    __del__ = close

    def DescriptorDistance(self, a, b):
        """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2015-2031); a, b: [..., 32] uint8 -> int32 [...]"""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.empty(a.shape[0], np.int32)
        check(lib().orbm_descriptor_distance(self._h, ptr(a), ptr(b), a.shape[0], ptr(out)))
        return out if out.size != 1 else int(out[0])

    # ---- guided searches (host arrays in, host arrays out; the frame / map points are given flattened, see capi.*_struct)
    def SearchByProjection(self, frame, mps, th=3.0, blocked=None, kp_to_mp=None):
        """ORBmatcher::SearchByProjection(pF, vpMapPoints, th) (src/ORBmatcher.cc:539-624).
        frame: dict(n_kp, kps_un, desc, bounds, scale_factors); mps: structured array capi.MP_DTYPE.
        -> (nmatches, kp_to_mp int32 [totalN]: index of the map point written into mvpMapPoints[g], -1 = untouched)"""
        fs, keep = capi.frame_struct(frame)
        mps = np.ascontiguousarray(mps, capi.MP_DTYPE)
        total = int(np.sum(keep["n_kp"]))
        blocked = np.zeros(total, np.uint8) if blocked is None else np.ascontiguousarray(blocked, np.uint8)
        out = np.full(total, -1, np.int32) if kp_to_mp is None else kp_to_mp
        n = C.c_int32()
        check(lib().orbm_search_by_projection(self._h, C.addressof(fs), ptr(mps), len(mps), th, self.mfNNratio, ptr(blocked), ptr(out), C.addressof(n)))
        return n.value, out

    def SearchByProjectionLast(self, cur, Rsw, tsw, K, last, th, bMapScaled=True, blocked=None):
        """ORBmatcher::SearchByProjection(pCurrentFrame, pLastFrame, th, bMapScaled) (src/ORBmatcher.cc:634-690, 954-1113).
        -> (nmatches, kp_to_last int32 [totalN], per_cam int32 [n_cams])"""
        fs, keep = capi.frame_struct(cur)
        ls, lkeep = capi.lastframe_struct(last)
        Rsw, tsw, K = (np.ascontiguousarray(a, np.float32) for a in (Rsw, tsw, K))
        total = int(np.sum(keep["n_kp"]))
        blocked = np.zeros(total, np.uint8) if blocked is None else np.ascontiguousarray(blocked, np.uint8)
        out = np.full(total, -1, np.int32)
        per_cam = np.zeros(len(keep["n_kp"]), np.int32)
        n = C.c_int32()
        check(lib().orbm_search_by_projection_last(self._h, C.addressof(fs), ptr(Rsw), ptr(tsw), ptr(K), C.addressof(ls), th, int(self.mbCheckOrientation),
                                                   int(bMapScaled), ptr(blocked), ptr(out), ptr(per_cam), C.addressof(n)))
        return n.value, out, per_cam

    def SearchByBoW(self, F, KF, kf_mp_valid, bMapScaled=True):
        """ORBmatcher::SearchByBoW(pF, pKF, vpMapPointMatches, bMapScaled) (src/ORBmatcher.cc:102-294).
        F, KF: dict(n_kp, desc, angle, node_first, node_id, node_off, idx) -> (nmatches, f_to_kf int32 [F totalN])"""
        fs, fk = capi.bowside_struct(F)
        ks, kk = capi.bowside_struct(KF)
        valid = np.ascontiguousarray(kf_mp_valid, np.uint8)
        out = np.full(int(np.sum(fk["n_kp"])), -1, np.int32)
        n = C.c_int32()
        check(lib().orbm_search_by_bow(self._h, C.addressof(fs), C.addressof(ks), ptr(valid), self.mfNNratio, int(self.mbCheckOrientation), int(bMapScaled),
                                       ptr(out), C.addressof(n)))
        return n.value, out

    def SearchByProjectionReloc(self, F, view, cam, points, th, ORBdist, blocked, kp_to_point=None):
        """ORBmatcher::SearchByProjectionOnCam(pF, query, pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:812-951).
        points: dict(valid, pos, max_dist, min_dist, desc, angle) over pKF's key points -> (nmatches, kp_to_point int32 [totalN])"""
        fs, fk = capi.frame_struct(F)
        vs, vk = capi.frustum_struct(view)
        ps, pk = capi.points_struct(points)
        blocked = np.ascontiguousarray(blocked, np.uint8)
        out = np.full(len(blocked), -1, np.int32) if kp_to_point is None else kp_to_point
        n = C.c_int32()
        check(lib().orbm_search_by_projection_reloc(self._h, C.addressof(fs), C.addressof(vs), int(cam), C.addressof(ps), float(th), int(ORBdist),
                                                    int(self.mbCheckOrientation), ptr(blocked), ptr(out), C.addressof(n)))
        return n.value, out

    def SearchByProjectionSim3(self, KF, view, cam, points, th, matched_local, kf_index_quirk=True):
        """ORBmatcher::SearchByProjection(pKF, query, Scq_w, vpPoints, vpMatched, th) (src/ORBmatcher.cc:416-536).
        view slot `cam` = the decomposed similarity -> (nmatches, local_to_point int32 [n_kp[cam]])"""
        fs, fk = capi.frame_struct(KF)
        vs, vk = capi.frustum_struct(view)
        ps, pk = capi.points_struct(points)
        matched_local = np.ascontiguousarray(matched_local, np.uint8)
        out = np.full(len(matched_local), -1, np.int32)
        n = C.c_int32()
        check(lib().orbm_search_by_projection_sim3(self._h, C.addressof(fs), C.addressof(vs), int(cam), C.addressof(ps), int(th), int(kf_index_quirk),
                                                   ptr(matched_local), ptr(out), C.addressof(n)))
        return n.value, out

    def ProjectBest(self, KF, view, points, th, variant, kf_index_quirk=True):
        """search part of SearchByProjection(pKF, vpMapPoints, sFound, th, ORBdist) / Fuse / Fuse(Scw) (src/ORBmatcher.cc:693-775, 1431-1527,
        1560-1668): variant in capi.KF_SEARCH / KF_FUSE / KF_FUSE_SIM3 -> (best_kp, best_dist) int32 [n_cams][n]"""
        fs, fk = capi.frame_struct(KF)
        vs, vk = capi.frustum_struct(view)
        ps, pk = capi.points_struct(points)
        shape = (len(fk["n_kp"]), len(pk["valid"]))
        bk = np.full(shape, -1, np.int32); bd = np.full(shape, 256, np.int32)
        check(lib().orbm_project_best(self._h, C.addressof(fs), C.addressof(vs), C.addressof(ps), float(th), int(variant), int(kf_index_quirk), ptr(bk), ptr(bd)))
        return bk, bd

    def ProjectBestCam(self, KF, view, points, th, variant, cam, kf_index_quirk=True):
        """one camera of ProjectBest -> (best_kp, best_dist) int32 [n]: the form for the reference's camera loops, which change the map
        points between cameras (src/ORBmatcher.cc:783-787, :1452)"""
        fs, fk = capi.frame_struct(KF)
        vs, vk = capi.frustum_struct(view)
        ps, pk = capi.points_struct(points)
        n = len(pk["valid"])
        bk = np.full(n, -1, np.int32); bd = np.full(n, 256, np.int32)
        check(lib().orbm_project_best_cam(self._h, C.addressof(fs), C.addressof(vs), C.addressof(ps), float(th), int(variant), int(kf_index_quirk), int(cam), ptr(bk), ptr(bd)))
        return bk, bd

    def SearchByBoWKF(self, K1, c1, K2, c2, mp_valid1, mp_valid2):
        """ORBmatcher::SearchByBoWCrossCam(pKF1, c1, pKF2, c2, vpMatches12) (src/ORBmatcher.cc:297-414)
        -> (nmatches, matches12 int32 [n_kp1[c1]] = global key point index in KF2 or -1)"""
        s1, k1 = capi.bowside_struct(K1)
        s2, k2 = capi.bowside_struct(K2)
        v1 = np.ascontiguousarray(mp_valid1, np.uint8); v2 = np.ascontiguousarray(mp_valid2, np.uint8)
        out = np.full(int(k1["n_kp"][c1]), -1, np.int32)
        n = C.c_int32()
        check(lib().orbm_search_by_bow_kf(self._h, C.addressof(s1), int(c1), C.addressof(s2), int(c2), ptr(v1), ptr(v2), self.mfNNratio,
                                          int(self.mbCheckOrientation), ptr(out), C.addressof(n)))
        return n.value, out

    def SearchForTriangulation(self, K1, K2, cam, kps1, kps2, has_mp1, has_mp2, F12, C1sw, R2sw, t2sw, K2cam, scale_factors):
        """ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, camS) (src/ORBmatcher.cc:1253-1427)
        -> (nmatches, matches12 int32 [n_kp1[cam]] = camera-local index in KF2 or -1)"""
        s1, k1 = capi.bowside_struct(K1)
        s2, k2 = capi.bowside_struct(K2)
        kps1 = np.ascontiguousarray(kps1, capi.KP_DTYPE); kps2 = np.ascontiguousarray(kps2, capi.KP_DTYPE)
        h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
        F12, C1sw, R2sw, t2sw, K2cam, sf = (np.ascontiguousarray(a, np.float32) for a in (F12, C1sw, R2sw, t2sw, K2cam, scale_factors))
        out = np.full(int(k1["n_kp"][cam]), -1, np.int32)
        n = C.c_int32()
        check(lib().orbm_search_for_triangulation(self._h, C.addressof(s1), C.addressof(s2), int(cam), ptr(kps1), ptr(kps2), ptr(h1), ptr(h2), ptr(F12),
                                                  ptr(C1sw), ptr(R2sw), ptr(t2sw), ptr(K2cam), ptr(sf), len(sf), int(self.mbCheckOrientation), ptr(out),
                                                  C.addressof(n)))
        return n.value, out

    def isInFrustum(self, frame, pos, normal, max_dist, min_dist, viewingCosLimit=0.5, bForAllCam=True):
        """Frame::isInFrustum + MapPoint::PredictScale for n map points (src/Frame.cc:244-312, src/MapPoint.cc:440-455).
        -> (out int32 [n][3] = in_view, cam, level ; uvc float32 [n][3] = u, v, viewCos)"""
        qs, keep = capi.frustum_struct(frame)
        pos, normal, max_dist, min_dist = (np.ascontiguousarray(a, np.float32) for a in (pos, normal, max_dist, min_dist))
        n = len(max_dist)
        out = np.zeros((n, 3), np.int32)
        uvc = np.zeros((n, 3), np.float32)
        check(lib().orbm_is_in_frustum(self._h, C.addressof(qs), ptr(pos), ptr(normal), ptr(max_dist), ptr(min_dist), n, viewingCosLimit, int(bForAllCam),
                                       ptr(out), ptr(uvc)))
        return out, uvc

    def UndistortKeyPoints(self, kps, K4, dist):
        """Frame::UndistortKeyPoints (src/Frame.cc:410-442): kps structured array (capi.KP_DTYPE) -> undistorted copy"""
        kps = np.ascontiguousarray(kps, capi.KP_DTYPE)
        K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        out = np.empty_like(kps)
        check(lib().orbm_undistort_keypoints(self._h, ptr(kps), len(kps), ptr(K4), ptr(dist), len(dist), ptr(out)))
        return out

    def ComputeImageBounds(self, width, height, K4, dist):
        """Frame::ComputeImageBounds (src/Frame.cc:454-490) -> float32 [4] = mvMinX, mvMaxX, mvMinY, mvMaxY"""
        K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        b = np.zeros(4, np.float32)
        check(lib().orbm_image_bounds(self._h, width, height, ptr(K4), ptr(dist), len(dist), ptr(b)))
        return b

    def bruteforce(self, dq, nq, dt, nt):
        """dq uint8 [P][Q][32], nq int32 [P], dt uint8 [P][T][32], nt int32 [P] -> best_idx, best_d, second_d int32 [P][Q]
        (entries >= nq[p] hold -1 / 256 / 256)."""
        dq = np.ascontiguousarray(dq, np.uint8)
        dt = np.ascontiguousarray(dt, np.uint8)
        nq = np.ascontiguousarray(nq, np.int32)
        nt = np.ascontiguousarray(nt, np.int32)
        P, Q, T = dq.shape[0], dq.shape[1], dt.shape[1]
        assert dt.shape[0] == P and nq.shape == (P,) and nt.shape == (P,)
        bi = np.full((P, Q), -1, np.int32)
        bd = np.full((P, Q), 256, np.int32)
        sd = np.full((P, Q), 256, np.int32)
        check(lib().orbm_bruteforce(self._h, ptr(dq), ptr(nq), Q, ptr(dt), ptr(nt), T, P, ptr(bi), ptr(bd), ptr(sd)))
        return bi, bd, sd

    def bruteforce_device(self, d_dq, d_nq, d_dt, d_nt, out=None, stream=None):
        """torch CUDA tensors; asynchronous on the current torch stream."""
        import torch
        P, Q, T = d_dq.shape[0], d_dq.shape[1], d_dt.shape[1]
        dev = d_dq.device
        if out is None:
            out = tuple(torch.empty((P, Q), dtype=torch.int32, device=dev) for _ in range(3))
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        L = lib()
        check(L.orbm_set_stream(self._h, _stream_handle(st)))
        check(L.orbm_bruteforce_device(self._h, d_dq.data_ptr(), d_nq.data_ptr(), Q, d_dt.data_ptr(), d_nt.data_ptr(), T, P,
                                       out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()))
        return out

    def bruteforce_sets_device(self, d_desc, d_counts, d_q_set, d_t_set, out=None, stream=None):
        """d_desc uint8 [S][cap][32] (any leading shape flattening to S sets), d_counts int32 [S], q/t set ids int32 [P]."""
        import torch
        cap = d_desc.shape[-2]
        S = d_desc.numel() // (cap * 32)
        P = d_q_set.numel()
        dev = d_desc.device
        if out is None:
            out = tuple(torch.empty((P, cap), dtype=torch.int32, device=dev) for _ in range(3))
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        L = lib()
        check(L.orbm_set_stream(self._h, _stream_handle(st)))
        check(L.orbm_bruteforce_sets_device(self._h, d_desc.data_ptr(), d_counts.data_ptr(), cap, S, d_q_set.data_ptr(), d_t_set.data_ptr(), P,
                                            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()))
        return out

    def profile(self, enable=True):
        check(lib().orbm_profile(self._h, int(enable)))

    def stage_ms(self):
        ms = C.c_double()
        n = C.c_int()
        check(lib().orbm_stage_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def synchronize(self):
        check(lib().orbm_synchronize(self._h))

    def launch_count(self):
        return int(lib().orbm_launch_count(self._h))
