"""Exact-arithmetic building blocks that the kernels and the oracle rely on bit for bit (orb_core.h, rbrief_pattern.h):
  * the rBRIEF test-pair table: product copy == oracle copy == the SHA-256 both headers quote == the reference's bit_pattern_31_
    (src/ORBextractor.cc:150-408) when the reference tree is present;
  * glibc_sincosf (the restatement of glibc 2.39 sincosf that steers the descriptor pattern, src/ORBextractor.cc:113) against this
    box's libm for EVERY float in [0, 2 pi + a margin] -- about 1.09e9 bit patterns, swept in C on all cores;
  * cv::fastAtan2 restatement at its quadrant boundaries."""
import ctypes as C
import hashlib
import os
import platform
import re
import threading

import numpy as np
import pytest

import model_lib as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERN_SHA = "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"


def _values(path, after):
    text = open(path).read()
    text = text[text.index(after):]
    body = text[text.index("{") + 1:text.index("}")] if "{" in text[:200] else text
    return np.array([int(v) for v in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/|//[^\n]*", "", body))], np.int64)


def test_rbrief_pattern_table():
    prod_txt = open(os.path.join(ROOT, "orb-slam2-dualcam_b200", "csrc", "rbrief_pattern.h")).read()
    body = prod_txt[prod_txt.index("#define ORB_RBRIEF_PATTERN_VALUES") + len("#define ORB_RBRIEF_PATTERN_VALUES"):]
    body = body[:body.index("static const")]                 # the macro's continuation lines
    prod = np.array([int(v) for v in re.findall(r"-?\d+", body)], np.int64)
    ora = _values(os.path.join(ROOT, "oracle", "rbrief_pattern_oracle.h"), "orb_oracle_pattern[1024]")
    ora = ora[1:] if len(ora) == 1025 else ora          # (the array size in the declarator)
    assert len(prod) == 1024 and len(ora) == 1024
    assert np.array_equal(prod, ora)
    assert prod.min() >= -13 and prod.max() <= 13
    assert hashlib.sha256(prod.astype(np.int8).tobytes()).hexdigest() == PATTERN_SHA
    assert PATTERN_SHA in prod_txt
    ref = "/root/reference/src/ORBextractor.cc"
    if os.path.exists(ref):
        txt = open(ref).read()
        blk = txt[txt.index("bit_pattern_31_[256*4]"):]
        blk = blk[blk.index("{") + 1:blk.index("};")]
        vals = np.array([int(v) for v in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", blk))], np.int64)
        assert len(vals) == 1024 and np.array_equal(vals, prod), "the table differs from the reference's bit_pattern_31_"


def test_sincosf_every_float_up_to_two_pi():
    if platform.libc_ver()[0] != "glibc":
        pytest.skip("the restatement targets glibc's sincosf")
    L = M.lib()
    L.hm_sincos_sweep.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    L.hm_sincos_sweep.restype = C.c_longlong
    hi = int(np.float32(6.2832).view(np.uint32))            # every angle the extractor can form: fastAtan2 degrees in [0, 360] x (float)(pi / 180)
    nt = os.cpu_count() or 4
    cuts = np.linspace(0, hi + 1, nt * 4 + 1).astype(np.int64)
    bad = [0] * (len(cuts) - 1)
    first = [None] * (len(cuts) - 1)

    def work(i):
        fb = C.c_uint32(0)
        bad[i] = L.hm_sincos_sweep(int(cuts[i]), int(cuts[i + 1] - 1), C.byref(fb))
        first[i] = fb.value

    pending = list(range(len(cuts) - 1))
    lock = threading.Lock()

    def runner():
        while True:
            with lock:
                if not pending:
                    return
                i = pending.pop()
            work(i)

    ts = [threading.Thread(target=runner) for _ in range(nt)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    total = sum(bad)
    if total:
        i = next(k for k, b in enumerate(bad) if b)
        x = np.uint32(first[i]).view(np.float32)
        pytest.fail(f"{total} floats in [0, 6.2832] differ from libm {platform.libc_ver()} (first: {x!r}, bits {first[i]:#x})")


def test_fast_atan2_quadrants():
    L = M.lib()
    assert L.hm_atan2(0.0, 1.0) == 0.0
    assert abs(L.hm_atan2(1.0, 0.0) - 90.0) < 1e-3
    assert abs(L.hm_atan2(0.0, -1.0) - 180.0) < 1e-3
    assert abs(L.hm_atan2(-1.0, 0.0) - 270.0) < 1e-3
    for y, x in [(3.0, 4.0), (-2.0, 7.0), (5.0, -1.0), (-6.0, -6.0)]:
        a = L.hm_atan2(y, x)
        t = np.degrees(np.arctan2(y, x)) % 360.0
        assert abs(a - t) < 0.3            # cv::fastAtan2 promises ~0.3 degrees
