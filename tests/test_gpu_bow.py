"""GPU parity (through the C-ABI) of the DBoW2 transform against oracle/bow_oracle.cpp: words, nodes, BowVector values (bit-exact doubles)
and the FeatureVector CSR; the FeatureVector then drives SearchByBoW end to end."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import ORBmatcher, ORBVocabulary, OrbError
import synth

pytestmark = pytest.mark.gpu


def _check_set(o, s, set_off, ref):
    lo, n = int(set_off[s]), int(set_off[s + 1] - set_off[s])
    assert np.array_equal(o["word_id"][lo:lo + n], ref["word_id"]) and np.array_equal(o["node_id"][lo:lo + n], ref["node_id"])
    nw, nf = int(o["n_words"][s]), int(o["n_fv_nodes"][s])
    assert nw == len(ref["bow_ids"]) and np.array_equal(o["bow_ids"][lo:lo + nw], ref["bow_ids"])
    assert o["bow_vals"][lo:lo + nw].tobytes() == ref["bow_vals"].tobytes()                  # bit-exact doubles
    assert nf == len(ref["fv_node"]) and np.array_equal(o["fv_node"][lo:lo + nf], ref["fv_node"])
    off = o["fv_off"][lo + s: lo + s + nf + 1]
    assert np.array_equal(off, ref["fv_off"])
    assert np.array_equal(o["fv_idx"][lo:lo + int(off[-1])], ref["fv_idx"])


@pytest.mark.parametrize("k,L,levelsup,ragged", [(10, 3, 1, False), (10, 4, 2, False), (6, 5, 4, True), (3, 2, 4, False)])
def test_transform_vs_oracle(k, L, levelsup, ragged):
    voc = synth.vocabulary(k * 10 + L, k=k, L=L, ragged=ragged)
    ns = [1000, 0, 1, 777, 2048, 13]
    set_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    desc = np.concatenate([synth.vocabulary_features(50 + i, voc, n) for i, n in enumerate(ns)])
    V = ORBVocabulary(voc)
    o = V.transform_batch(desc, set_off, levelsup)
    R = O.Vocabulary(voc)
    for s in range(len(ns)):
        _check_set(o, s, set_off, R.transform(desc[set_off[s]:set_off[s + 1]], levelsup))
    assert V.size() == int(voc["is_leaf"].sum())


def test_largest_set_and_limits():
    voc = synth.vocabulary(9, k=10, L=3)
    desc = synth.vocabulary_features(10, voc, 8192)
    V = ORBVocabulary(voc)
    o = V.transform_batch(desc, [0, 8192], 2)
    _check_set(o, 0, np.array([0, 8192]), O.Vocabulary(voc).transform(desc, 2))
    with pytest.raises(OrbError):
        V.transform_batch(np.zeros((8193, 32), np.uint8), [0, 8193], 2)
    bad = dict(voc, parent=voc["parent"].copy())
    bad["parent"][5] = 7                                    # a child listed before its parent
    with pytest.raises(OrbError):
        ORBVocabulary(bad)


def test_feature_vectors_drive_search_by_bow():
    """Frame::ComputeBoW -> ORBmatcher::SearchByBoW with device-made feature vectors equals the oracle chain"""
    voc = synth.vocabulary(21, k=10, L=4)
    V, R = ORBVocabulary(voc), O.Vocabulary(voc)
    kf_desc = [synth.vocabulary_features(30 + c, voc, n) for c, n in enumerate((900, 800))]
    f_desc = [synth.random_descriptors(40 + c, len(d), 0.04, d[np.random.default_rng(c).permutation(len(d))]) for c, d in enumerate(kf_desc)]
    rng = np.random.default_rng(5)

    def side(descs, fv):
        n = sum(len(d) for d in descs)
        return dict(n_kp=np.array([len(d) for d in descs], np.int32), desc=np.concatenate(descs), angle=rng.uniform(0, 360, n).astype(np.float32), **fv)

    def ref_fv(descs):
        node_first, node_id, node_off, idx = [0], [], [0], []
        for d in descs:
            r = R.transform(d, 2)
            node_id.extend(r["fv_node"].tolist()); base = len(idx); idx.extend(r["fv_idx"].tolist()); node_off.extend((base + r["fv_off"][1:]).tolist())
            node_first.append(len(node_id))
        return dict(node_first=np.array(node_first, np.int32), node_id=np.array(node_id, np.int32), node_off=np.array(node_off, np.int32), idx=np.array(idx, np.int32))

    (_, fvK, _), (_, fvF, _) = V.transform(kf_desc, 2), V.transform(f_desc, 2)
    rK, rF = ref_fv(kf_desc), ref_fv(f_desc)
    for k in rK:
        assert np.array_equal(fvK[k], rK[k]) and np.array_equal(fvF[k], rF[k]), k
    KF, F = side(kf_desc, fvK), side(f_desc, fvF)
    valid = (rng.random(len(KF["desc"])) < 0.8).astype(np.uint8)
    n, out = ORBmatcher(0.7, False).SearchByBoW(F, KF, valid)
    rn, rout = O.search_by_bow(F, KF, valid, 0.7, False, True)
    assert n == rn and np.array_equal(out, rout) and n > 300
