"""The GPU formulation of each extractor stage (orb_core.h / orb_geometry.h, emulated serially on the CPU by
tests/model/host_model.cpp) against the oracle.  Runs without a GPU."""
import numpy as np
import pytest

import model_lib as M
import oracle_lib as O
import synth


def _images():
    rng = np.random.default_rng(11)
    yield "textured", synth.dual_sequence(0, 1, 640, 480, cams=1)[0, 0], 1000
    yield "textured720p", synth.dual_sequence(1, 1, 1280, 720, cams=1)[0, 0], 2000
    yield "noise", rng.integers(0, 256, (240, 320), dtype=np.uint8), 500
    yield "lowcontrast", (synth.dual_sequence(3, 1, 320, 240, cams=1)[0, 0] // 8 + 100).astype(np.uint8), 500
    yield "flat", np.full((200, 300), 77, np.uint8), 300
    yield "plateaus", np.kron(rng.integers(0, 2, (30, 40), dtype=np.uint8) * 200, np.ones((8, 8), np.uint8)), 400


@pytest.mark.parametrize("name,img,nf", list(_images()), ids=[n for n, _, _ in _images()])
def test_fast_cells_and_quadtree_formulation(name, img, nf):
    H, W = img.shape
    ex = O.Extractor(nfeatures=nf)
    ex(img)
    geo, max_kp = M.geometry(W, H, nfeatures=nf)
    total = 0
    for l in range(8):
        g = geo[l]
        lw, lh = ex.level_size(l)
        assert (g["w"], g["h"]) == (lw, lh)
        ref_c = ex.level_candidates(l)
        lvl = ex.level_pixels(l)
        got = M.fast_level(lvl)
        got_sorted = M.unpack(M.sort_reference_order(got, g)) if len(got) else np.zeros((0, 3), np.int32)
        assert np.array_equal(got_sorted, ref_c), f"{name} level {l}: FAST candidates differ"
        # quadtree on a shuffled candidate list: the GPU list is unordered
        ref_s = ex.level_selected(l)
        if len(got) == 0:
            assert len(ref_s) == 0
            continue
        rng = np.random.default_rng(l)
        sel = M.unpack(M.quadtree(got[rng.permutation(len(got))], g))
        sel[:, 0] += 16
        sel[:, 1] += 16
        assert np.array_equal(sel, ref_s), f"{name} level {l}: quadtree differs"
        total += len(sel)
    assert total <= max_kp


def test_quadtree_random_point_sets():
    """Random sparse/dense/clustered point sets straight into both quadtrees (stress on tie-breaks and the sorted phase)."""
    import ctypes as C
    L = O.lib()
    rng = np.random.default_rng(5)
    geo, _ = M.geometry(640, 480)
    for trial in range(60):
        g = dict(geo[rng.integers(0, 8)])
        g["quota"] = int(rng.integers(1, 300))
        n = int(rng.integers(1, 3000))
        if trial % 3 == 0:   # clustered
            cx, cy = rng.integers(3, g["width"] - 3), rng.integers(3, g["height"] - 3)
            xs = np.clip(cx + rng.integers(-12, 13, n), 3, g["width"] - 4)
            ys = np.clip(cy + rng.integers(-12, 13, n), 3, g["height"] - 4)
        else:
            xs = rng.integers(3, g["width"] - 3, n)
            ys = rng.integers(3, g["height"] - 3, n)
        pts = np.unique(np.stack([xs, ys], 1), axis=0)
        sc = rng.integers(7, 40 if trial % 2 else 255, len(pts))
        cands = M.pack(np.concatenate([pts, sc[:, None]], 1))
        ref_sorted = M.sort_reference_order(cands, g)
        got = M.unpack(M.quadtree(cands[rng.permutation(len(cands))], g))
        # oracle quadtree through a private hook: feed candidates via a synthetic extractor call is not possible,
        # so compare against the python list implementation used for the golden files
        from tools_make_golden import distribute
        p = [tuple(r) for r in M.unpack(ref_sorted).tolist()]
        keep = distribute(p, 16, 16 + g["width"], 16, 16 + g["height"], g["quota"])
        ref = np.array([p[k] for k in keep], np.int32).reshape(-1, 3)
        assert np.array_equal(got, ref), f"trial {trial}"


def test_sincos_matches_glibc_dense_sample():
    hi = np.float32(6.2832).view(np.uint32)
    bits = np.arange(0, int(hi) + 1, 251, dtype=np.uint32)
    x = bits.view(np.float32)
    c0, s0 = O.cosf_sinf(x)
    c1, s1 = M.sincos(x)
    assert np.array_equal(c0.view(np.uint32), c1.view(np.uint32))
    assert np.array_equal(s0.view(np.uint32), s1.view(np.uint32))


def test_atan2_matches_oracle():
    rng = np.random.default_rng(2)
    Lm, Lo = M.lib(), O.lib()
    for y, x in rng.integers(-70000, 70000, (20000, 2)).astype(np.float32).tolist() + [[0, 0], [0, -1], [-1, 0], [1, 0], [0, 1]]:
        assert np.float32(Lm.hm_atan2(y, x)).view(np.uint32) == np.float32(Lo.orc_fast_atan2(y, x)).view(np.uint32)
