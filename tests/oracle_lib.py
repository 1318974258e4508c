"""ctypes binding of the CPU oracle (oracle/_build/liborb_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def build():
    """(Re)build the oracle shared object with the recipe committed under oracle/."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "_build", "liborb_oracle.so")
    odir = os.path.join(ROOT, "oracle")
    srcs = [os.path.join(odir, f) for f in os.listdir(odir) if f.endswith((".cpp", ".h"))]
    if (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build()
    L = C.CDLL(path)
    L.orc_resize_linear_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
    L.orc_gaussian7_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
    L.orc_fast9.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
    L.orc_fast9.restype = C.c_int
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_fast_atan2.restype = C.c_float
    L.orc_cvround_f.argtypes = [C.c_float]
    L.orc_cvround_f.restype = C.c_int
    L.orc_hamming256.argtypes = [u8p, u8p]
    L.orc_hamming256.restype = C.c_int
    L.orc_cosf_sinf.argtypes = [f32p, C.c_int, f32p, f32p]
    L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.orc_extractor_create.restype = C.c_void_p
    L.orc_extractor_destroy.argtypes = [C.c_void_p]
    L.orc_extractor_nlevels.argtypes = [C.c_void_p]
    L.orc_extractor_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
    L.orc_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, u8p, C.c_int]
    L.orc_extract.restype = C.c_int
    L.orc_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_level_pixels.argtypes = [C.c_void_p, C.c_int]
    L.orc_level_pixels.restype = C.c_void_p
    L.orc_level_blurred.argtypes = [C.c_void_p, C.c_int]
    L.orc_level_blurred.restype = C.c_void_p
    L.orc_level_candidates.argtypes = [C.c_void_p, C.c_int, i32p, C.c_int]
    L.orc_level_candidates.restype = C.c_int
    L.orc_level_selected.argtypes = [C.c_void_p, C.c_int, i32p, C.c_int]
    L.orc_level_selected.restype = C.c_int
    L.orc_extractor_timers.argtypes = [C.c_void_p, f64p]
    L.orc_match_bruteforce.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p, i32p]
    _LIB = L
    return L


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(src, u8p), src.shape[1], src.shape[0], src.strides[0], _ptr(dst, u8p), dw, dh, dw)
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian7_u8(_ptr(src, u8p), src.shape[1], src.shape[0], src.strides[0], _ptr(dst, u8p), dst.strides[0])
    return dst


def fast9(img, threshold, nonmax=True, cap=1 << 16):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9(_ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], threshold, int(nonmax), _ptr(out, i32p), cap)
    assert n <= cap
    return out[:n].copy()


def cosf_sinf(x):
    x = np.ascontiguousarray(x, np.float32)
    c = np.empty_like(x)
    s = np.empty_like(x)
    lib().orc_cosf_sinf(_ptr(x, f32p), x.size, _ptr(c, f32p), _ptr(s, f32p))
    return c, s


class Extractor:
    """Mirror of ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) on the oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        assert self.h
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        fpl = np.empty(n, np.int32)
        umax = np.empty(16, np.int32)
        self.L.orc_extractor_tables(self.h, _ptr(sc, f32p), _ptr(isc, f32p), _ptr(s2, f32p), _ptr(is2, f32p), _ptr(fpl, i32p), _ptr(umax, i32p))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, features_per_level=fpl, umax=umax)

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _ptr(img, u8p), img.shape[1], img.shape[0], img.strides[0], kps.ctypes.data, _ptr(desc, u8p), cap)
        assert 0 <= n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_pixels(self, l, blurred=False):
        w, h = self.level_size(l)
        p = (self.L.orc_level_blurred if blurred else self.L.orc_level_pixels)(self.h, l)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(h, w)).copy()

    def level_candidates(self, l, cap=1 << 17):
        out = np.empty((cap, 3), np.int32)
        n = self.L.orc_level_candidates(self.h, l, _ptr(out, i32p), cap)
        return out[:min(n, cap)].copy()

    def level_selected(self, l, cap=1 << 14):
        out = np.empty((cap, 3), np.int32)
        n = self.L.orc_level_selected(self.h, l, _ptr(out, i32p), cap)
        return out[:min(n, cap)].copy()

    def timers(self):
        t = np.zeros(6, np.float64)
        self.L.orc_extractor_timers(self.h, _ptr(t, f64p))
        return dict(zip(["pyramid", "fast", "quadtree", "orient", "blur", "desc"], t.tolist()))


def match_bruteforce(dq, dt):
    dq = np.ascontiguousarray(dq, np.uint8)
    dt = np.ascontiguousarray(dt, np.uint8)
    nq, nt = dq.shape[0], dt.shape[0]
    bi, bd, sd = (np.empty(nq, np.int32) for _ in range(3))
    lib().orc_match_bruteforce(_ptr(dq, u8p), nq, _ptr(dt, u8p), nt, _ptr(bi, i32p), _ptr(bd, i32p), _ptr(sd, i32p))
    return bi, bd, sd


# ------------------------------------------------------------------------------------------------ bundle adjustment
class BAProblemC(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_cams", C.c_int32),
                ("poses", C.c_void_p), ("pose_fixed", C.c_void_p), ("points", C.c_void_p),
                ("edge_pose", C.c_void_p), ("edge_point", C.c_void_p), ("edge_cam", C.c_void_p),
                ("edge_obs", C.c_void_p), ("edge_inv_sigma2", C.c_void_p),
                ("cam_K", C.c_void_p), ("cam_ext", C.c_void_p), ("cam_adj", C.c_void_p)]


class BAStatsC(C.Structure):
    _fields_ = [("initial_chi2", C.c_double), ("final_chi2", C.c_double), ("final_lambda", C.c_double),
                ("iterations", C.c_int32), ("trials", C.c_int32), ("outliers", C.c_int32)]


BA_KEYS = [("poses", np.float64), ("pose_fixed", np.uint8), ("points", np.float64), ("edge_pose", np.int32), ("edge_point", np.int32),
           ("edge_cam", np.int32), ("edge_obs", np.float64), ("edge_inv_sigma2", np.float64), ("cam_K", np.float64),
           ("cam_ext", np.float64), ("cam_adj", np.float64)]


def ba_struct(p, cls=BAProblemC):
    """dict of numpy arrays (synth.ba_problem layout) -> (C struct, keep-alive list)"""
    keep = {k: np.ascontiguousarray(p[k], dt) for k, dt in BA_KEYS}
    s = cls(n_poses=keep["pose_fixed"].shape[0], n_points=keep["points"].shape[0], n_edges=keep["edge_pose"].shape[0],
            n_cams=keep["cam_K"].shape[0], **{k: keep[k].ctypes.data for k, _ in BA_KEYS})
    return s, keep


HUBER_MONO = float(np.float32(np.sqrt(5.991)))   # `const float thHuberMono = sqrt(5.991)`  src/Optimizer.cc:514
HUBER_2D = float(np.float32(np.sqrt(3.99)))       # `const float thHuber2D = sqrt(3.99)`  src/Optimizer.cc:108 (BundleAdjustment)


def local_ba(p, its1=5, its2=10, huber_delta=HUBER_MONO, chi2_th=5.991, stop=None):
    L = lib()
    L.orc_local_ba.argtypes = [C.POINTER(BAProblemC), C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BAStatsC)]
    s, keep = ba_struct(p)
    poses = np.zeros((s.n_poses, 12)); points = np.zeros((s.n_points, 3)); out = np.zeros(s.n_edges, np.uint8)
    st = BAStatsC()
    rc = L.orc_local_ba(C.byref(s), its1, its2, huber_delta, chi2_th, stop.ctypes.data if stop is not None else None,
                        poses.ctypes.data, points.ctypes.data, out.ctypes.data, C.byref(st))
    return rc, poses, points, out.astype(bool), {f: getattr(st, f) for f, _ in BAStatsC._fields_}


def global_ba(p, iterations=10, huber_delta=HUBER_2D, stop=None):
    L = lib()
    L.orc_global_ba.argtypes = [C.POINTER(BAProblemC), C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BAStatsC)]
    s, keep = ba_struct(p)
    poses = np.zeros((s.n_poses, 12)); points = np.zeros((s.n_points, 3))
    st = BAStatsC()
    rc = L.orc_global_ba(C.byref(s), iterations, huber_delta, stop.ctypes.data if stop is not None else None,
                         poses.ctypes.data, points.ctypes.data, C.byref(st))
    return rc, poses, points, {f: getattr(st, f) for f, _ in BAStatsC._fields_}


def ba_normal_equations(p, huber_delta=HUBER_MONO):
    L = lib()
    L.orc_ba_normal_equations.argtypes = [C.POINTER(BAProblemC), C.c_double] + [C.c_void_p] * 6
    s, keep = ba_struct(p)
    K = int((keep["pose_fixed"] == 0).sum())
    Hpp = np.zeros((K, 6, 6)); bp = np.zeros((K, 6)); Hll = np.zeros((s.n_points, 3, 3)); bl = np.zeros((s.n_points, 3))
    Hpl = np.zeros((s.n_edges, 6, 3)); chi2 = C.c_double()
    k = L.orc_ba_normal_equations(C.byref(s), huber_delta, Hpp.ctypes.data, bp.ctypes.data, Hll.ctypes.data, bl.ctypes.data,
                                  Hpl.ctypes.data, C.addressof(chi2))
    assert k == K
    return Hpp, bp, Hll, bl, Hpl, chi2.value


# ------------------------------------------------------------------------------------------------ guided searches
def _capi():
    from orbslam2_dualcam_b200 import capi
    return capi


def search_by_projection(frame, mps, th=3.0, nnratio=0.6, blocked=None):
    cp = _capi()
    fs, keep = cp.frame_struct(frame)
    mps = np.ascontiguousarray(mps, cp.MP_DTYPE)
    total = int(np.sum(keep["n_kp"]))
    blocked = np.zeros(total, np.uint8) if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    out = np.full(total, -1, np.int32)
    L = lib()
    L.orc_search_by_projection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection.restype = C.c_int
    n = L.orc_search_by_projection(C.addressof(fs), mps.ctypes.data, len(mps), th, nnratio, blocked.ctypes.data, out.ctypes.data)
    return n, out


def search_by_projection_last(cur, Rsw, tsw, K, last, th, check_ori=True, map_scaled=True, blocked=None):
    cp = _capi()
    fs, keep = cp.frame_struct(cur)
    ls, lkeep = cp.lastframe_struct(last)
    Rsw, tsw, K = (np.ascontiguousarray(a, np.float32) for a in (Rsw, tsw, K))
    total = int(np.sum(keep["n_kp"]))
    blocked = np.zeros(total, np.uint8) if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    out = np.full(total, -1, np.int32)
    per_cam = np.zeros(len(keep["n_kp"]), np.int32)
    L = lib()
    L.orc_search_by_projection_last.argtypes = [C.c_void_p] * 5 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_last.restype = C.c_int
    n = L.orc_search_by_projection_last(C.addressof(fs), Rsw.ctypes.data, tsw.ctypes.data, K.ctypes.data, C.addressof(ls), th, int(check_ori), int(map_scaled),
                                        blocked.ctypes.data, out.ctypes.data, per_cam.ctypes.data)
    return n, out, per_cam


def search_by_bow(F, KF, kf_mp_valid, nnratio=0.7, check_ori=True, map_scaled=True):
    cp = _capi()
    fs, fk = cp.bowside_struct(F)
    ks, kk = cp.bowside_struct(KF)
    valid = np.ascontiguousarray(kf_mp_valid, np.uint8)
    out = np.full(int(np.sum(fk["n_kp"])), -1, np.int32)
    L = lib()
    L.orc_search_by_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
    L.orc_search_by_bow.restype = C.c_int
    n = L.orc_search_by_bow(C.addressof(fs), C.addressof(ks), valid.ctypes.data, nnratio, int(check_ori), int(map_scaled), out.ctypes.data)
    return n, out


def is_in_frustum(frame, pos, normal, max_dist, min_dist, cos_limit=0.5, for_all=True):
    cp = _capi()
    qs, keep = cp.frustum_struct(frame)
    pos, normal, max_dist, min_dist = (np.ascontiguousarray(a, np.float32) for a in (pos, normal, max_dist, min_dist))
    n = len(max_dist)
    out = np.zeros((n, 3), np.int32)
    uvc = np.zeros((n, 3), np.float32)
    L = lib()
    L.orc_is_in_frustum.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_is_in_frustum.restype = None
    L.orc_is_in_frustum(C.addressof(qs), pos.ctypes.data, normal.ctypes.data, max_dist.ctypes.data, min_dist.ctypes.data, n, cos_limit, int(for_all),
                        out.ctypes.data, uvc.ctypes.data)
    return out, uvc


def pose_optimization(f):
    """Optimizer::PoseOptimization on the oracle -> (pose [12], outlier bool [n], n_inliers, (iterations, trials))"""
    L = lib()
    L.orc_pose_optimization.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 6
    L.orc_pose_optimization.restype = C.c_int
    k = dict(pose=np.ascontiguousarray(f["pose"], np.float64).reshape(12), Xw=np.ascontiguousarray(f["Xw"], np.float64), obs=np.ascontiguousarray(f["obs"], np.float64),
             inv_sigma2=np.ascontiguousarray(f["inv_sigma2"], np.float64), cam=np.ascontiguousarray(f["cam"], np.int32),
             cam_K=np.ascontiguousarray(f["cam_K"], np.float64), cam_ext=np.ascontiguousarray(f["cam_ext"], np.float64), cam_adj=np.ascontiguousarray(f["cam_adj"], np.float64))
    n = len(k["inv_sigma2"])
    pose = np.zeros(12); out = np.zeros(max(n, 1), np.uint8); cnt = np.zeros(2, np.int32)
    r = L.orc_pose_optimization(k["pose"].ctypes.data, n, k["Xw"].ctypes.data, k["obs"].ctypes.data, k["inv_sigma2"].ctypes.data, k["cam"].ctypes.data,
                                k["cam_K"].shape[0], k["cam_K"].ctypes.data, k["cam_ext"].ctypes.data, k["cam_adj"].ctypes.data, pose.ctypes.data,
                                out.ctypes.data, cnt.ctypes.data)
    return pose, out[:n].astype(bool), r, (int(cnt[0]), int(cnt[1]))


def undistort_points(pts, K4, dist):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(pts)
    L = lib()
    L.orc_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_undistort_points.restype = None
    L.orc_undistort_points(pts.ctypes.data, len(pts), K4.ctypes.data, dist.ctypes.data, len(dist), out.ctypes.data)
    return out


def image_bounds(width, height, K4, dist):
    K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    L = lib()
    L.orc_image_bounds.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_image_bounds.restype = None
    L.orc_image_bounds(width, height, K4.ctypes.data, dist.ctypes.data, len(dist), b.ctypes.data)
    return b


# ---- key-frame flavoured searches (oracle/match_kf_oracle.cpp)
def search_by_projection_reloc(F, view, cam, points, th, ORBdist, blocked, check_ori=True):
    cp = _capi()
    fs, fk = cp.frame_struct(F); vs, vk = cp.frustum_struct(view); ps, pk = cp.points_struct(points)
    blocked = np.ascontiguousarray(blocked, np.uint8)
    out = np.full(len(blocked), -1, np.int32)
    L = lib()
    L.orc_search_by_projection_reloc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_reloc.restype = C.c_int
    n = L.orc_search_by_projection_reloc(C.addressof(fs), C.addressof(vs), cam, C.addressof(ps), th, ORBdist, int(check_ori), blocked.ctypes.data, out.ctypes.data)
    return n, out


def search_by_projection_sim3(KF, view, cam, points, th, matched_local, kf_quirk=True):
    cp = _capi()
    fs, fk = cp.frame_struct(KF); vs, vk = cp.frustum_struct(view); ps, pk = cp.points_struct(points)
    matched_local = np.ascontiguousarray(matched_local, np.uint8)
    out = np.full(len(matched_local), -1, np.int32)
    L = lib()
    L.orc_search_by_projection_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_search_by_projection_sim3.restype = C.c_int
    n = L.orc_search_by_projection_sim3(C.addressof(fs), C.addressof(vs), cam, C.addressof(ps), th, int(kf_quirk), matched_local.ctypes.data, out.ctypes.data)
    return n, out


def project_best(KF, view, points, th, variant, kf_quirk=True):
    """variant 0: SearchByProjection(KF, MPs, ...) ; 1: Fuse ; 2: Fuse(Scw)"""
    cp = _capi()
    fs, fk = cp.frame_struct(KF); vs, vk = cp.frustum_struct(view); ps, pk = cp.points_struct(points)
    shape = (len(fk["n_kp"]), len(pk["valid"]))
    bk = np.zeros(shape, np.int32); bd = np.zeros(shape, np.int32)
    L = lib()
    if variant == 0:
        L.orc_search_kf_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_search_kf_points.restype = None
        L.orc_search_kf_points(C.addressof(fs), C.addressof(vs), C.addressof(ps), th, int(kf_quirk), bk.ctypes.data, bd.ctypes.data)
    else:
        L.orc_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_fuse.restype = None
        L.orc_fuse(C.addressof(fs), C.addressof(vs), C.addressof(ps), th, int(variant == 2), int(kf_quirk), bk.ctypes.data, bd.ctypes.data)
    return bk, bd


def search_by_bow_kf(K1, c1, K2, c2, mp_valid1, mp_valid2, nnratio=0.7, check_ori=True):
    cp = _capi()
    s1, k1 = cp.bowside_struct(K1); s2, k2 = cp.bowside_struct(K2)
    v1 = np.ascontiguousarray(mp_valid1, np.uint8); v2 = np.ascontiguousarray(mp_valid2, np.uint8)
    out = np.full(int(k1["n_kp"][c1]), -1, np.int32)
    L = lib()
    L.orc_search_by_bow_kf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    L.orc_search_by_bow_kf.restype = C.c_int
    n = L.orc_search_by_bow_kf(C.addressof(s1), c1, C.addressof(s2), c2, v1.ctypes.data, v2.ctypes.data, nnratio, int(check_ori), out.ctypes.data)
    return n, out


def search_for_triangulation(K1, K2, cam, kps1, kps2, has_mp1, has_mp2, F12, C1sw, R2sw, t2sw, K2cam, scale_factors, check_ori=True):
    cp = _capi()
    s1, k1 = cp.bowside_struct(K1); s2, k2 = cp.bowside_struct(K2)
    kps1 = np.ascontiguousarray(kps1, cp.KP_DTYPE); kps2 = np.ascontiguousarray(kps2, cp.KP_DTYPE)
    h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
    F12, C1sw, R2sw, t2sw, K2cam, sf = (np.ascontiguousarray(a, np.float32) for a in (F12, C1sw, R2sw, t2sw, K2cam, scale_factors))
    out = np.full(int(k1["n_kp"][cam]), -1, np.int32)
    L = lib()
    L.orc_search_for_triangulation.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_int, C.c_void_p]
    L.orc_search_for_triangulation.restype = C.c_int
    n = L.orc_search_for_triangulation(C.addressof(s1), C.addressof(s2), cam, kps1.ctypes.data, kps2.ctypes.data, h1.ctypes.data, h2.ctypes.data, F12.ctypes.data,
                                       C1sw.ctypes.data, R2sw.ctypes.data, t2sw.ctypes.data, K2cam.ctypes.data, sf.ctypes.data, int(check_ori), out.ctypes.data)
    return n, out


# ---- DBoW2 transform (oracle/bow_oracle.cpp)
class Vocabulary:
    def __init__(self, voc):
        L = lib()
        L.orc_vocab_create.restype = C.c_void_p
        L.orc_vocab_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._k = [np.ascontiguousarray(voc["parent"], np.int32), np.ascontiguousarray(voc["is_leaf"], np.uint8), np.ascontiguousarray(voc["desc"], np.uint8),
                   np.ascontiguousarray(voc["weight"], np.float64)]
        self._h = L.orc_vocab_create(int(voc["k"]), int(voc["L"]), len(self._k[0]), *[a.ctypes.data for a in self._k])
        assert self._h, "bad vocabulary"

    def __del__(self):
        L = lib()
        L.orc_vocab_destroy.argtypes = [C.c_void_p]
        L.orc_vocab_destroy.restype = None
        if getattr(self, "_h", None):
            L.orc_vocab_destroy(self._h)
            self._h = None

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); ids = np.zeros(n, np.int32); vals = np.zeros(n, np.float64)
        fvn = np.zeros(n, np.int32); fvo = np.zeros(n + 1, np.int32); fvi = np.zeros(n, np.int32)
        nw = C.c_int32(); nf = C.c_int32()
        L = lib()
        L.orc_vocab_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9
        L.orc_vocab_transform.restype = C.c_int
        L.orc_vocab_transform(self._h, desc.ctypes.data, n, levelsup, word.ctypes.data, node.ctypes.data, ids.ctypes.data, vals.ctypes.data, C.addressof(nw),
                              fvn.ctypes.data, fvo.ctypes.data, fvi.ctypes.data, C.addressof(nf))
        return dict(word_id=word, node_id=node, bow_ids=ids[:nw.value], bow_vals=vals[:nw.value], fv_node=fvn[:nf.value], fv_off=fvo[:nf.value + 1],
                    fv_idx=fvi[:fvo[nf.value]])


def bow_score_l1(a, b):
    L = lib()
    L.orc_bow_score_l1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_bow_score_l1.restype = C.c_double
    i1, v1 = np.ascontiguousarray(a[0], np.int32), np.ascontiguousarray(a[1], np.float64)
    i2, v2 = np.ascontiguousarray(b[0], np.int32), np.ascontiguousarray(b[1], np.float64)
    return L.orc_bow_score_l1(i1.ctypes.data, v1.ctypes.data, len(i1), i2.ctypes.data, v2.ctypes.data, len(i2))
