"""Pins the oracle's restatements of the OpenCV primitives against the cv2 wheel in this container
(cv2 4.13.0: the only runnable OpenCV; the reference's OpenCV is an un-vendored system dependency)."""
import numpy as np
import pytest

import oracle_lib as O
import synth

cv2 = pytest.importorskip("cv2")


def _img(seed, w=640, h=480):
    return synth.dual_sequence(seed, 1, w, h, cams=1)[0, 0]


@pytest.mark.parametrize("seed", [0, 1])
def test_resize_chain_matches_cv2(seed):
    img = _img(seed)
    ex = O.Extractor()
    inv = ex.tables()["inv_scale"]
    prev = img
    for l in range(1, 8):
        w = int(np.rint(np.float32(640) * inv[l]))
        h = int(np.rint(np.float32(480) * inv[l]))
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        got = O.resize_linear(prev, w, h)
        assert np.array_equal(ref, got), f"level {l}"
        prev = ref


@pytest.mark.parametrize("size", [(37, 23), (101, 77), (640, 480), (1280, 720)])
def test_resize_random_sizes(size):
    rng = np.random.default_rng(size[0])
    src = rng.integers(0, 256, (size[1], size[0]), dtype=np.uint8)
    for f in (1.2, 1.44, 1.07, 2.5):
        w, h = max(int(round(size[0] / f)), 1), max(int(round(size[1] / f)), 1)
        assert np.array_equal(cv2.resize(src, (w, h), interpolation=cv2.INTER_LINEAR), O.resize_linear(src, w, h))


@pytest.mark.parametrize("size", [(8, 9), (64, 48), (179, 134), (640, 480)])
def test_gaussian7_matches_cv2(size):
    rng = np.random.default_rng(size[0] * 7 + 1)
    src = rng.integers(0, 256, (size[1], size[0]), dtype=np.uint8)
    ref = cv2.GaussianBlur(src, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, O.gaussian7(src))
    smooth = _img(3, max(size[0], 16), max(size[1], 16))
    assert np.array_equal(cv2.GaussianBlur(smooth, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101), O.gaussian7(smooth))


@pytest.mark.parametrize("th", [7, 20, 0, 60])
def test_fast9_matches_cv2(th):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rng = np.random.default_rng(th)
    imgs = [_img(5)[100:140, 200:243], _img(6)[0:36, 0:36], rng.integers(0, 256, (41, 38), dtype=np.uint8),
            rng.integers(100, 140, (40, 43), dtype=np.uint8), _img(7)[:120, :160], np.full((20, 20), 9, np.uint8)]
    for im in imgs:
        im = np.ascontiguousarray(im)
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(im)]
        got = [tuple(r) for r in O.fast9(im, th).tolist()]
        assert ref == got


def test_fast9_no_nms_matches_cv2():
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=False, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    im = np.ascontiguousarray(_img(8)[:100, :100])
    ref = [(int(k.pt[0]), int(k.pt[1])) for k in det.detect(im)]
    got = [tuple(r[:2]) for r in O.fast9(im, 20, nonmax=False).tolist()]
    assert ref == got


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = np.concatenate([rng.integers(-60000, 60000, 30000), [0, 0, 1, -1, 0, 5, -5]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-60000, 60000, 30000), [0, 1, 0, 0, -1, 5, -5]]).astype(np.float32)
    L = O.lib()
    for y, x in zip(ys.tolist(), xs.tolist()):
        assert np.float32(cv2.fastAtan2(y, x)) == np.float32(L.orc_fast_atan2(y, x)), (y, x)


def test_cvround_half_even():
    L = O.lib()
    for v, r in [(0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (2.4999, 2), (-2.5001, -3), (13.0, 13)]:
        assert L.orc_cvround_f(v) == r
