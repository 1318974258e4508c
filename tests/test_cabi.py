"""The C-ABI library loads on a CPU-only box and exports every symbol include/orbslam2_dualcam_b200.h declares;
without a GPU every create() fails loudly with ORB_E_NO_DEVICE (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import orbslam2_dualcam_b200 as orb
from orbslam2_dualcam_b200 import capi
import synth

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "orbslam2_dualcam_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[xmbv]?a?_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    L = orb.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature in capi.py"
    for n in capi.SIGNATURES:
        assert n in names, f"{n} bound in capi.py but not declared in the header"


def test_version_and_error_text():
    L = orb.lib()
    assert b"sm_100a" in L.orb_version()
    assert isinstance(L.orb_last_error(), bytes)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="needs a box without a GPU")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(orb.OrbError) as e:
        orb.ORBextractor()
    assert e.value.code == capi.ORB_E_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(orb.OrbError):
        orb.ORBmatcher()
    with pytest.raises(orb.OrbError):
        orb.Optimizer()
    with pytest.raises(orb.OrbError) as e:
        orb.ORBVocabulary(synth.vocabulary(0, k=3, L=2))
    assert e.value.code == capi.ORB_E_NO_DEVICE


def test_null_handles_are_rejected():
    L = orb.lib()
    assert L.orbx_synchronize(None) == capi.ORB_E_INVALID
    assert L.orbm_synchronize(None) == capi.ORB_E_INVALID
    assert L.orbba_synchronize(None) == capi.ORB_E_INVALID
    assert L.orbx_launch_count(None) == 0
    h = C.c_void_p()
    assert L.orbx_create(C.byref(h), 0, 10, 10, 1, 1, 1000, 1.2, 8, 20, 7) == capi.ORB_E_INVALID   # argument check comes first
    assert b"image size" in L.orb_last_error()
