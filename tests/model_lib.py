"""ctypes binding of tests/model/host_model.cpp: CPU emulation of the GPU formulation (test support)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
u8p, u32p, i32p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_float))


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    src = os.path.join(ROOT, "tests", "model", "host_model.cpp")
    out = os.path.join(ROOT, "tests", "model", "_build", "libhost_model.so")
    deps = [src] + [os.path.join(ROOT, "orb-slam2-dualcam_b200", "csrc", f) for f in ("orb_core.h", "orb_geometry.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.hm_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, i32p, f32p, C.POINTER(C.c_int)]
    L.hm_fast_level.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u32p, C.c_int]
    L.hm_fast_level.restype = C.c_int
    L.hm_order_key.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
    L.hm_order_key.restype = C.c_uint32
    L.hm_quadtree.argtypes = [u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p, C.c_int]
    L.hm_quadtree.restype = C.c_int
    L.hm_sincos.argtypes = [f32p, C.c_int, f32p, f32p]
    L.hm_atan2.argtypes = [C.c_float, C.c_float]
    L.hm_atan2.restype = C.c_float
    _LIB = L
    return L


GEO_FIELDS = ["w", "h", "width", "height", "nCols", "nRows", "wCell", "hCell", "nColsEff", "nRowsEff", "nIni", "quota", "valid"]


def geometry(W, H, nfeatures=1000, sf=1.2, nlevels=8, ini=20, mn=7):
    o = np.zeros((nlevels, 13), np.int32)
    hx = np.zeros(nlevels, np.float32)
    mk = C.c_int()
    lib().hm_geometry(W, H, nfeatures, sf, nlevels, ini, mn, o.ctypes.data_as(i32p), hx.ctypes.data_as(f32p), C.byref(mk))
    lv = [dict(zip(GEO_FIELDS, row.tolist()), hX=float(h)) for row, h in zip(o, hx)]
    return lv, mk.value


def unpack(c):
    c = np.asarray(c, np.uint32)
    return np.stack([(c & 0xfff).astype(np.int32), ((c >> 12) & 0xfff).astype(np.int32), (c >> 24).astype(np.int32)], 1)


def pack(xys):
    xys = np.asarray(xys, np.uint32)
    return (xys[:, 0] | (xys[:, 1] << 12) | (xys[:, 2] << 24)).astype(np.uint32)


def fast_level(img, ini=20, mn=7, cap=1 << 17):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(cap, np.uint32)
    n = lib().hm_fast_level(img.ctypes.data_as(u8p), img.shape[1], img.shape[0], ini, mn, out.ctypes.data_as(u32p), cap)
    assert n <= cap
    return out[:n].copy()


def order_keys(cands, g):
    L = lib()
    return np.array([L.hm_order_key(int(c), g["wCell"], g["hCell"], g["nColsEff"], g["nRowsEff"]) for c in cands], np.uint32)


def sort_reference_order(cands, g):
    return cands[np.argsort(order_keys(cands, g), kind="stable")]


def quadtree(cands, g, cap=4096):
    cands = np.ascontiguousarray(cands, np.uint32)
    out = np.zeros(cap, np.uint32)
    n = lib().hm_quadtree(cands.ctypes.data_as(u32p), len(cands), g["width"], g["height"], g["nIni"], g["hX"], g["quota"],
                          g["wCell"], g["hCell"], g["nColsEff"], g["nRowsEff"], out.ctypes.data_as(u32p), cap)
    assert 0 <= n <= cap, n
    return out[:n].copy()


def sincos(x):
    x = np.ascontiguousarray(x, np.float32)
    c, s = np.empty_like(x), np.empty_like(x)
    lib().hm_sincos(x.ctypes.data_as(f32p), x.size, c.ctypes.data_as(f32p), s.ctypes.data_as(f32p))
    return c, s
