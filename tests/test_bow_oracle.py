"""CPU checks of the DBoW2 transform oracle (oracle/bow_oracle.cpp) against a direct numpy restatement on a small vocabulary, and of the
host-side vocabulary text parser."""
import numpy as np

import oracle_lib as O
from orbslam2_dualcam_b200 import parse_text_vocabulary
import synth


def _numpy_transform(voc, desc, levelsup):
    children = {}
    for nid in range(1, len(voc["parent"])):
        children.setdefault(int(voc["parent"][nid]), []).append(nid)
    words = np.cumsum(voc["is_leaf"]) - 1
    bow, fv, word, node = {}, {}, [], []
    for i, f in enumerate(desc):
        cur, level, nid = 0, 0, 0
        while cur in children:
            level += 1
            ch = children[cur]
            d = [int(np.unpackbits(f ^ voc["desc"][c]).sum()) for c in ch]
            cur = ch[int(np.argmin(d))]               # argmin returns the first minimum
            if level == voc["L"] - levelsup:
                nid = cur
        word.append(int(words[cur])); node.append(nid)
        w = float(voc["weight"][cur])
        if w > 0:
            bow[words[cur]] = bow.get(words[cur], 0.0) + w
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    norm = 0.0
    for k in ids:
        norm += abs(bow[k])
    return np.array(word), np.array(node), np.array(ids), np.array([bow[k] / norm for k in ids]), fv


def test_transform_matches_numpy_restatement():
    voc = synth.vocabulary(1, k=5, L=3)
    desc = synth.vocabulary_features(2, voc, 300)
    V = O.Vocabulary(voc)
    for levelsup in (1, 2, 3, 4):
        r = V.transform(desc, levelsup)
        word, node, ids, vals, fv = _numpy_transform(voc, desc, levelsup)
        assert np.array_equal(r["word_id"], word) and np.array_equal(r["node_id"], node)
        assert np.array_equal(r["bow_ids"], ids) and np.array_equal(r["bow_vals"], vals)            # bit-exact: same summation order
        assert abs(r["bow_vals"].sum() - 1.0) < 1e-12
        assert list(r["fv_node"]) == sorted(fv)
        for j, nid in enumerate(r["fv_node"]):
            assert list(r["fv_idx"][r["fv_off"][j]:r["fv_off"][j + 1]]) == fv[nid]
    assert (V.transform(desc, 3)["node_id"] == 0).all()                # L - levelsup <= 0: every feature hangs off the root


def test_stopped_words_and_score():
    voc = synth.vocabulary(3, k=4, L=2, frac_stopped=0.3)
    V = O.Vocabulary(voc)
    a = V.transform(synth.vocabulary_features(4, voc, 200), 1)
    stopped = voc["weight"][np.flatnonzero(voc["is_leaf"])][a["word_id"]] == 0
    assert stopped.any() and len(a["fv_idx"]) == (~stopped).sum()
    b = V.transform(synth.vocabulary_features(5, voc, 150), 1)
    sa, sb = (a["bow_ids"], a["bow_vals"]), (b["bow_ids"], b["bow_vals"])
    assert abs(O.bow_score_l1(sa, sa) - 1.0) < 1e-12 and 0.0 <= O.bow_score_l1(sa, sb) < 1.0
    assert O.bow_score_l1(sa, sb) == O.bow_score_l1(sb, sa)


def test_text_vocabulary_round_trip():
    voc = synth.vocabulary(6, k=3, L=3, ragged=True)
    back = parse_text_vocabulary(synth.vocabulary_text(voc))
    for k in ("parent", "is_leaf", "desc", "weight"):
        assert np.array_equal(back[k], voc[k]), k
    assert back["k"] == 3 and back["L"] == 3
