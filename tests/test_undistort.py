"""Frame::UndistortKeyPoints / ComputeImageBounds: the oracle restatement of cv::undistortPoints is pinned bit-exactly to the committed
cv2 golden vectors (and to the live cv2 where it is installed); the GPU kernel is bit-exact against the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort.npz"))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("i", range(4))
def test_oracle_matches_cv2_golden(i):
    got = O.undistort_points(G["pts"], G["K4"], G[f"dist{i}"])
    assert np.array_equal(_bits(got), _bits(G[f"und{i}"]))


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    pts = rng.uniform([0, 0], [1280, 720], (5000, 2)).astype(np.float32)
    K4 = np.array([900.0, 905.0, 640.5, 355.2], np.float32)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    d = np.array([-0.21, 0.05, 3e-4, -2e-4, 0.001], np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
    assert np.array_equal(_bits(O.undistort_points(pts, K4, d)), _bits(ref))


def test_image_bounds_oracle():
    b = O.image_bounds(640, 480, G["K4"], G["dist0"])
    und = G["und0"][-4:]          # the four image corners are the last golden points
    assert b[0] == min(und[0, 0], und[2, 0]) and b[1] == max(und[1, 0], und[3, 0]) and b[2] == min(und[0, 1], und[1, 1]) and b[3] == max(und[2, 1], und[3, 1])
    assert np.array_equal(O.image_bounds(640, 480, G["K4"], np.zeros(5, np.float32)), np.array([0, 640, 0, 480], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(4))
def test_gpu_undistort_keypoints(i):
    from orbslam2_dualcam_b200 import ORBmatcher, capi
    m = ORBmatcher()
    kps = np.zeros(len(G["pts"]), capi.KP_DTYPE)
    kps["x"] = G["pts"][:, 0]; kps["y"] = G["pts"][:, 1]; kps["octave"] = np.arange(len(kps)) % 8; kps["angle"] = 12.5; kps["size"] = 31; kps["response"] = 40
    un = m.UndistortKeyPoints(kps, G["K4"], G[f"dist{i}"])
    assert np.array_equal(_bits(np.stack([un["x"], un["y"]], 1)), _bits(G[f"und{i}"]))
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(un[f], kps[f])
    b = m.ComputeImageBounds(640, 480, G["K4"], G[f"dist{i}"])
    assert np.array_equal(_bits(b), _bits(O.image_bounds(640, 480, G["K4"], G[f"dist{i}"])))


@pytest.mark.gpu
def test_gpu_undistort_no_distortion_is_a_copy():
    from orbslam2_dualcam_b200 import ORBmatcher, capi
    m = ORBmatcher()
    kps = np.zeros(10, capi.KP_DTYPE)
    kps["x"] = np.arange(10) * 3.5; kps["y"] = 7
    assert m.UndistortKeyPoints(kps, G["K4"], np.zeros(5, np.float32)).tobytes() == kps.tobytes()      # src/Frame.cc:414-418
    assert np.array_equal(m.ComputeImageBounds(640, 480, G["K4"], np.zeros(4, np.float32)), np.array([0, 640, 0, 480], np.float32))
