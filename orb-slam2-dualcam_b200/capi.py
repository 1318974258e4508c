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
    "orbba_synchronize": (C.c_int, [vp]),
    "orbba_launch_count": (C.c_longlong, [vp]),
    "orbba_local": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp, vp]),
    "orbba_global": (C.c_int, [vp, vp, C.c_int, C.c_double, vp, vp, vp, vp]),
    "orbba_upload": (C.c_int, [vp, vp, C.c_int]),
    "orbba_run": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_double]),
    "orbba_download": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
    "orbba_download_batch": (C.c_int, [vp, vp, vp, vp, vp]),
    "orbba_kernel_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbba_profile": (C.c_int, [vp, C.c_int]),
    "orbba_stage_ms": (C.c_int, [vp, f64p, C.POINTER(C.c_int)]),
    "orbm_bruteforce_sets_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp]),
}

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
