"""The drop-in adaptor sources (adaptor/*.cc: bodies of ORBextractor / ORBmatcher / Optimizer on top of the C-ABI) and the native
harnesses.  CPU part: every adaptor source passes a compiler front end against the shim headers (the extractor also against the
reference's own include/ORBextractor.h when /root/reference is present) and the harnesses link against the library.  GPU part: the
harnesses run -- the hot path called natively from C and through the adaptor's ORB_SLAM2::ORBextractor, no Python in between."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTOR = os.path.join(ROOT, "adaptor")


def _make(target=None):
    from orbslam2_dualcam_b200 import build
    build.build()
    cmd = ["make", "-s", "-C", ADAPTOR] + ([target] if target else [])
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_adaptor_sources_compile():
    r = _make("check")
    assert r.returncode == 0, r.stdout + r.stderr


def test_harnesses_link():
    r = _make()
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("harness_c", "harness_extractor"):
        assert os.path.exists(os.path.join(ADAPTOR, "_build", name))


@pytest.mark.gpu
@pytest.mark.parametrize("name,needle", [("harness_c", "harness ok"), ("harness_extractor", "identical to orbx_extract: yes")])
def test_harness_runs(name, needle):
    exe = os.path.join(ADAPTOR, "_build", name)
    if not os.path.exists(exe):
        r = _make()
        assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and needle in r.stdout, r.stdout + r.stderr
