#!/usr/bin/env python
"""Generates tests/golden/extract_*.npz: golden outputs of ORBextractor::operator() produced by an INDEPENDENT
composition that calls the real OpenCV primitives (cv2 wheel) exactly where the reference calls them:

    cv2.resize INTER_LINEAR            <- ComputePyramid            (src/ORBextractor.cc:1120)
    cv2 FAST(9/16, nms) per 30-px cell <- ComputeKeyPointsOctTree   (src/ORBextractor.cc:789-829)
    python list quadtree               <- DistributeOctTree         (src/ORBextractor.cc:539-763)
    cv2.fastAtan2                      <- IC_Angle                  (src/ORBextractor.cc:77-104)
    cv2.GaussianBlur(7,7,2)            <- operator()                (src/ORBextractor.cc:1085-1086)
    glibc cosf/sinf + float32 numpy    <- computeOrbDescriptor      (src/ORBextractor.cc:108-147)

It needs cv2 and is run in the build container only; the .npz fixtures it writes are committed and are what
tests/test_oracle_golden.py (CPU) and tests/test_gpu_extract.py (GPU) compare against.
Determinisation of the address tie-break is the same as the oracle's (creation order), see oracle/orb_oracle.cpp.
"""
import ctypes
import math
import os
import re
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth

f32 = np.float32
libm = ctypes.CDLL("libm.so.6")
libm.cosf.restype = ctypes.c_float
libm.cosf.argtypes = [ctypes.c_float]
libm.sinf.restype = ctypes.c_float
libm.sinf.argtypes = [ctypes.c_float]


def load_pattern():
    txt = open(os.path.join(ROOT, "orb-slam2-dualcam_b200", "csrc", "rbrief_pattern.h")).read()
    body = txt[txt.index("#define ORB_RBRIEF_PATTERN_VALUES"):txt.index("static const int8_t")]
    body = body.split("\n", 1)[1]
    nums = [int(x) for x in re.findall(r"-?\d+", body)]
    assert len(nums) == 1024
    return np.array(nums, np.int32).reshape(256, 4)


def cv_round(v):
    return int(np.rint(v))


from py_quadtree import distribute  # noqa: E402


def extract(img, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniTh=20, minTh=7):
    pattern = load_pattern()
    scale = [f32(1.0)]
    for i in range(1, nlevels):
        scale.append(f32(scale[-1] * f32(scaleFactor)))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0) / f32(scaleFactor)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    per = []
    for _ in range(nlevels - 1):
        per.append(cv_round(nd)); nd = f32(nd * factor)
    per.append(max(nfeatures - sum(per), 0))
    umax = [0] * 16
    vmax = int(math.floor(15 * math.sqrt(2.0) / 2 + 1)); vmin = int(math.ceil(15 * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(225.0 - v * v))
    v0 = 0
    for v in range(15, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0; v0 += 1
    H, Wd = img.shape
    pyr = [img]
    for l in range(1, nlevels):
        sz = (cv_round(f32(Wd) * inv[l]), cv_round(f32(H) * inv[l]))
        pyr.append(cv2.resize(pyr[l - 1], sz, interpolation=cv2.INTER_LINEAR))
    det_ini = cv2.FastFeatureDetector_create(threshold=iniTh, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(threshold=minTh, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps_all, desc_all = [], []
    for l in range(nlevels):
        im = pyr[l]
        minBX = minBY = 16
        maxBX, maxBY = im.shape[1] - 16, im.shape[0] - 16
        width, height = f32(maxBX - minBX), f32(maxBY - minBY)
        nCols, nRows = int(width / f32(30)), int(height / f32(30))
        wCell, hCell = int(math.ceil(width / f32(nCols))), int(math.ceil(height / f32(nRows)))
        pts = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
                k = det_ini.detect(roi)
                if not k:
                    k = det_min.detect(roi)
                for kp in k:
                    pts.append((int(kp.pt[0]) + j * wCell, int(kp.pt[1]) + i * hCell, int(kp.response)))
        keep = distribute(pts, minBX, maxBX, minBY, maxBY, per[l])
        if not keep:
            continue
        blurred = cv2.GaussianBlur(im.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        size = float(int(f32(31) * scale[l]))
        for k in keep:
            x, y, resp = pts[k][0] + minBX, pts[k][1] + minBY, pts[k][2]
            m01 = m10 = 0
            for u in range(-15, 16):
                m10 += u * int(im[y, x + u])
            for v in range(1, 16):
                d = umax[v]
                plus = im[y + v, x - d:x + d + 1].astype(np.int64); minus = im[y - v, x - d:x + d + 1].astype(np.int64)
                us = np.arange(-d, d + 1)
                m01 += v * int((plus - minus).sum())
                m10 += int((us * (plus + minus)).sum())
            angle = f32(cv2.fastAtan2(float(m01), float(m10)))
            rad = f32(angle * f32(math.pi / 180.0))
            a, b = f32(libm.cosf(float(rad))), f32(libm.sinf(float(rad)))
            px = pattern[:, [0, 2]].astype(np.float32); py = pattern[:, [1, 3]].astype(np.float32)
            rr = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
            cc = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
            vals = blurred[y + rr, x + cc]
            bits = (vals[:, 0] < vals[:, 1]).astype(np.uint8)
            desc_all.append(np.packbits(bits, bitorder="little"))
            sx, sy = (f32(x), f32(y)) if l == 0 else (f32(x) * scale[l], f32(y) * scale[l])
            kps_all.append((sx, sy, size, angle, float(resp), l, -1))
    kps = np.array(kps_all, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
    desc = np.array(desc_all, np.uint8).reshape(-1, 32)
    return kps, desc


CASES = {
    # name: (image factory, extractor args)
    "textured_640x480": (lambda: synth.dual_sequence(0, 1, 640, 480, cams=1)[0, 0], dict()),
    "lowcontrast_320x240": (lambda: (synth.dual_sequence(3, 1, 320, 240, cams=1)[0, 0] // 8 + 100).astype(np.uint8), dict(nfeatures=500)),
    "sparse_400x300": (lambda: np.ascontiguousarray(np.pad(synth.dual_sequence(5, 1, 120, 90, cams=1)[0, 0], ((100, 110), (140, 140)), constant_values=50)), dict(nfeatures=300)),
}


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, (mk, args) in CASES.items():
        img = mk()
        kps, desc = extract(img, **args)
        np.savez_compressed(os.path.join(out, f"extract_{name}.npz"), img=img, kps=kps, desc=desc,
                            args=np.array([args.get("nfeatures", 1000), 8, 20, 7], np.int32))
        print(name, img.shape, len(kps), "keypoints")


if __name__ == "__main__":
    main()
