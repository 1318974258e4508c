"""ctypes binding of the C-ABI library (include/orbslam2_dualcam_b200.h).  No compute happens in Python and there is
no fallback: if the library is missing, `lib()` raises."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "liborbslam2_dualcam_b200.so")

ORB_OK, ORB_E_INVALID, ORB_E_NO_DEVICE, ORB_E_CUDA, ORB_E_OVERFLOW, ORB_E_ABORTED = 0, -1, -2, -3, -4, -5

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

u8p, i32p, f32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_float, C.c_double))
vp = C.c_void_p

# every symbol include/orbslam2_dualcam_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "orb_last_error": (C.c_char_p, []),
    "orb_version": (C.c_char_p, []),
    "orbx_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]),
    "orbx_destroy": (None, [vp]),
    "orbx_get_levels": (C.c_int, [vp]),
    "orbx_get_tables": (C.c_int, [vp, f32p, f32p, f32p, f32p, i32p, i32p]),
    "orbx_max_keypoints": (C.c_int, [vp]),
    "orbx_level_size": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "orbx_extract": (C.c_int, [vp, vp, C.c_int, C.c_size_t, vp, vp, vp, C.c_int]),
    "orbx_extract_device": (C.c_int, [vp, vp, C.c_int, C.c_size_t, vp, vp, vp, C.c_int]),
    "orbx_set_stream": (C.c_int, [vp, vp]),
    "orbx_synchronize": (C.c_int, [vp]),
    "orbx_launch_count": (C.c_longlong, [vp]),
    "orbx_profile": (C.c_int, [vp, C.c_int]),
    "orbx_stage_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbx_debug_level": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
    "orbx_debug_candidates": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int]),
    "orbx_debug_selected": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int]),
    "orbm_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]),
    "orbm_destroy": (None, [vp]),
    "orbm_set_stream": (C.c_int, [vp, vp]),
    "orbm_synchronize": (C.c_int, [vp]),
    "orbm_launch_count": (C.c_longlong, [vp]),
    "orbm_profile": (C.c_int, [vp, C.c_int]),
    "orbm_stage_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbm_descriptor_distance": (C.c_int, [vp, vp, vp, C.c_int, vp]),
    "orbm_bruteforce": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "orbm_bruteforce_device": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "orbba_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
    "orbba_destroy": (None, [vp]),
    "orbba_set_stream": (C.c_int, [vp, vp]),
    "orbba_set_copy_stream": (C.c_int, [vp, vp]),
    "orbba_synchronize": (C.c_int, [vp]),
    "orbba_launch_count": (C.c_longlong, [vp]),
    "orbba_local": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp, vp]),
    "orbba_local_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp, vp]),
    "orbba_global": (C.c_int, [vp, vp, C.c_int, C.c_double, vp, vp, vp, vp]),
    "orbba_upload": (C.c_int, [vp, vp, C.c_int]),
    "orbba_upload_f32": (C.c_int, [vp, vp, C.c_int]),
    "orbba_run": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_double]),
    "orbba_download": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
    "orbba_download_batch": (C.c_int, [vp, vp, vp, vp, vp]),
    "orbba_kernel_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbba_profile": (C.c_int, [vp, C.c_int]),
    "orbba_stage_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbm_search_by_projection": (C.c_int, [vp, vp, vp, C.c_int, C.c_float, C.c_float, vp, vp, vp]),
    "orbm_search_by_projection_last": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, vp, vp, vp]),
    "orbm_search_by_bow": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, vp]),
    "orbm_is_in_frustum": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_int, vp, vp]),
    "orbba_pose_optimization": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp]),
    "orbba_dist_unique_id": (C.c_int, [vp]),
    "orbba_dist_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]),
    "orbba_dist_destroy": (None, [vp]),
    "orbba_dist_launch_count": (C.c_longlong, [vp]),
    "orbba_dist_optimize": (C.c_int, [vp, vp, C.c_int, C.c_double, vp, vp, vp, vp]),
    "orbba_dist_timing": (C.c_int, [vp, f64p, f64p, f64p]),
    "orbba_dist_loop_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_longlong)]),
    "orbba_dist_segments": (C.c_int, [vp]),
    "orbm_search_by_projection_reloc": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_float, C.c_int, C.c_int, vp, vp, vp]),
    "orbm_search_by_projection_sim3": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp]),
    "orbm_project_best": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, vp]),
    "orbm_project_best_cam": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, C.c_int, vp, vp]),
    "orbm_search_by_bow_kf": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_float, C.c_int, vp, vp]),
    "orbm_search_for_triangulation": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "orbv_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "orbv_destroy": (None, [vp]),
    "orbv_set_stream": (C.c_int, [vp, vp]),
    "orbv_words": (C.c_int, [vp]),
    "orbv_launch_count": (C.c_longlong, [vp]),
    "orbv_transform": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "orbm_undistort_keypoints": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "orbm_image_bounds": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "orbm_bruteforce_sets_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp]),
}

# ---- PODs of the guided searches (include/orbslam2_dualcam_b200.h)
MP_DTYPE = np.dtype([("valid", "<i4"), ("cam", "<i4"), ("u", "<f4"), ("v", "<f4"), ("level", "<i4"), ("view_cos", "<f4"),
                     ("obs_positive", "<i4"), ("desc", "u1", (32,))])
assert MP_DTYPE.itemsize == 60


class FrameC(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_kp", vp), ("kps_un", vp), ("desc", vp), ("bounds", vp), ("n_levels", C.c_int32), ("scale_factors", vp)]


class LastFrameC(C.Structure):
    _fields_ = [("n", C.c_int32), ("cam", vp), ("valid", vp), ("pos", vp), ("desc", vp), ("octave", vp), ("angle", vp), ("obs_positive", vp)]


class BowSideC(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_kp", vp), ("desc", vp), ("angle", vp), ("node_first", vp), ("node_id", vp), ("node_off", vp), ("idx", vp)]


class FrustumC(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_levels", C.c_int32), ("Rsw", vp), ("tsw", vp), ("Ow", vp), ("K", vp), ("bounds", vp),
                ("log_scale_factor", C.c_float)]


KF_SEARCH, KF_FUSE, KF_FUSE_SIM3 = 0, 1, 2     # ORBM_KF_* variants of orbm_project_best


class PointsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("valid", vp), ("pos", vp), ("normal", vp), ("max_dist", vp), ("min_dist", vp), ("desc", vp), ("angle", vp)]


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def frame_struct(f, cls=FrameC):
    """dict(n_kp, kps_un (KP_DTYPE), desc, bounds, scale_factors) -> (struct, keep-alive)"""
    keep = dict(n_kp=_c(f["n_kp"], np.int32), kps_un=_c(f["kps_un"], KP_DTYPE), desc=_c(f["desc"], np.uint8), bounds=_c(f["bounds"], np.float32),
                scale_factors=_c(f["scale_factors"], np.float32))
    s = cls(n_cams=len(keep["n_kp"]), n_kp=keep["n_kp"].ctypes.data, kps_un=keep["kps_un"].ctypes.data, desc=keep["desc"].ctypes.data,
            bounds=keep["bounds"].ctypes.data, n_levels=len(keep["scale_factors"]), scale_factors=keep["scale_factors"].ctypes.data)
    return s, keep


def lastframe_struct(l, cls=LastFrameC):
    keep = dict(cam=_c(l["cam"], np.int32), valid=_c(l["valid"], np.uint8), pos=_c(l["pos"], np.float32), desc=_c(l["desc"], np.uint8),
                octave=_c(l["octave"], np.int32), angle=_c(l["angle"], np.float32), obs_positive=_c(l["obs_positive"], np.uint8))
    s = cls(n=len(keep["cam"]), **{k: v.ctypes.data for k, v in keep.items()})
    return s, keep


def bowside_struct(b, cls=BowSideC):
    keep = dict(n_kp=_c(b["n_kp"], np.int32), desc=_c(b["desc"], np.uint8), angle=_c(b["angle"], np.float32), node_first=_c(b["node_first"], np.int32),
                node_id=_c(b["node_id"], np.int32), node_off=_c(b["node_off"], np.int32), idx=_c(b["idx"], np.int32))
    s = cls(n_cams=len(keep["n_kp"]), **{k: v.ctypes.data for k, v in keep.items()})
    return s, keep


def frustum_struct(q, cls=FrustumC):
    keep = dict(Rsw=_c(q["Rsw"], np.float32), tsw=_c(q["tsw"], np.float32), Ow=_c(q["Ow"], np.float32), K=_c(q["K"], np.float32),
                bounds=_c(q["bounds"], np.float32))
    s = cls(n_cams=keep["Rsw"].shape[0], n_levels=int(q["n_levels"]), log_scale_factor=float(q["log_scale_factor"]), **{k: v.ctypes.data for k, v in keep.items()})
    return s, keep


def points_struct(p, cls=PointsC):
    """dict(valid, pos, normal (opt), max_dist, min_dist, desc, angle (opt)) -> (struct, keep-alive)"""
    keep = dict(valid=_c(p["valid"], np.uint8), pos=_c(p["pos"], np.float32), max_dist=_c(p["max_dist"], np.float32), min_dist=_c(p["min_dist"], np.float32),
                desc=_c(p["desc"], np.uint8))
    for k in ("normal", "angle"):
        if p.get(k) is not None:
            keep[k] = _c(p[k], np.float32)
    s = cls(n=len(keep["valid"]), **{k: v.ctypes.data for k, v in keep.items()})
    return s, keep


_LIB = None


class OrbError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"orbslam2_dualcam_b200 error {code}: {text}")
        self.code = code


def lib():
    """Load the C-ABI library.  Raises if it has not been built: there is no CPU or eager fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                          "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(L, name):
            continue   # optional groups (search / BA) are bound by their own modules once built
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def check(rc):
    if rc < 0:
        raise OrbError(rc, lib().orb_last_error().decode("utf-8", "replace"))
    return rc


def ptr(a):
    """address of a numpy array / torch tensor / int"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def stream_handle(torch_stream):
    """cudaStream_t value for the C-ABI.  torch's default stream is the legacy NULL stream, whose handle 0 means "the
    handle's own stream" in orbx_set_stream / orbm_set_stream; cudaStreamLegacy (0x1) names it explicitly."""
    return torch_stream.cuda_stream or 1
