"""Oracle extractor vs the committed golden vectors (made by tools/make_golden.py from real cv2 primitives)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "extract_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_oracle_extract_matches_golden(path):
    g = np.load(path)
    nf = int(g["args"][0])
    ex = O.Extractor(nfeatures=nf)
    kps, desc = ex(g["img"])
    assert len(kps) == len(g["kps"])
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], g["kps"][f]), f
    assert np.array_equal(desc, g["desc"])


def test_golden_present():
    assert len(GOLDEN) >= 3


def test_extractor_tables_known_answers():
    t = O.Extractor().tables()
    # SURVEY.md §8 a1 known answers
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert t["features_per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    t2 = O.Extractor(nfeatures=2000).tables()
    assert t2["features_per_level"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
