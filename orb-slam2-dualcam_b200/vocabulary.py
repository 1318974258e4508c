"""Host mirror of ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for the one call the path makes:
transform(features, BowVector, FeatureVector, levelsup) from Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:393-408).
All computation is in orb_bow.cu behind orbv_*; this module parses the vocabulary text file and shapes the outputs."""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, ptr


def parse_text_vocabulary(lines):
    """TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1362-1447): header "k L scoring weighting",
    then one line per node: parent, leaf flag, 32 descriptor bytes, weight.  -> dict(k, L, parent, is_leaf, desc, weight) with row 0 = root."""
    it = iter(lines)
    head = next(it).split()
    k, L, n1, n2 = int(head[0]), int(head[1]), int(head[2]), int(head[3])
    if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
        raise ValueError("Vocabulary loading failure: This is not a correct text file!")
    if (n1, n2) != (0, 0):
        raise ValueError("only the ORB vocabulary's L1_NORM scoring / TF_IDF weighting (header '... 0 0') is implemented")
    parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
    for line in it:
        tok = line.split()
        if not tok:
            continue
        parent.append(int(tok[0])); leaf.append(int(tok[1]) > 0)
        desc.append(np.array(tok[2:34], np.int64).astype(np.uint8)); weight.append(float(tok[34]))
    return dict(k=k, L=L, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8), desc=np.stack(desc), weight=np.array(weight, np.float64))


class ORBVocabulary:
    def __init__(self, voc, device=0):
        """voc: dict(k, L, parent, is_leaf, desc, weight) (see parse_text_vocabulary)"""
        self._h = C.c_void_p()
        self.k, self.L = int(voc["k"]), int(voc["L"])
        self._keep = dict(parent=np.ascontiguousarray(voc["parent"], np.int32), is_leaf=np.ascontiguousarray(voc["is_leaf"], np.uint8),
                          desc=np.ascontiguousarray(voc["desc"], np.uint8), weight=np.ascontiguousarray(voc["weight"], np.float64))
        k = self._keep
        check(lib().orbv_create(C.byref(self._h), device, self.k, self.L, len(k["parent"]), ptr(k["parent"]), ptr(k["is_leaf"]), ptr(k["desc"]), ptr(k["weight"])))

    @classmethod
    def loadFromTextFile(cls, filename, device=0):
        with open(filename) as f:
            return cls(parse_text_vocabulary(f), device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().orbv_destroy(h)

    def size(self):
        return lib().orbv_words(self._h)

    @property
    def launches(self):
        return lib().orbv_launch_count(self._h)

    def transform_batch(self, desc, set_off, levelsup=4):
        """desc uint8 [n][32], set_off int32 [n_sets + 1] -> raw arrays of orbv_transform (see the header)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        set_off = np.ascontiguousarray(set_off, np.int32)
        n, S = len(desc), len(set_off) - 1
        o = dict(word_id=np.full(n, -1, np.int32), node_id=np.full(n, -1, np.int32), bow_ids=np.full(n, -1, np.int32), bow_vals=np.zeros(n, np.float64),
                 n_words=np.zeros(S, np.int32), fv_node=np.full(n, -1, np.int32), fv_off=np.zeros(n + S, np.int32), fv_idx=np.full(n, -1, np.int32),
                 n_fv_nodes=np.zeros(S, np.int32))
        check(lib().orbv_transform(self._h, ptr(desc), ptr(set_off), S, int(levelsup), ptr(o["word_id"]), ptr(o["node_id"]), ptr(o["bow_ids"]), ptr(o["bow_vals"]),
                                   ptr(o["n_words"]), ptr(o["fv_node"]), ptr(o["fv_off"]), ptr(o["fv_idx"]), ptr(o["n_fv_nodes"])))
        return o

    def transform(self, desc_per_cam, levelsup=4):
        """Frame::ComputeBoW for one frame: list of per-camera descriptor arrays -> (bow, fv, raw): bow = per camera (ids, values);
        fv = the orbm_bowside_t CSR fields (node_first, node_id, node_off, idx) that ORBmatcher.SearchByBoW* take."""
        ns = [len(d) for d in desc_per_cam]
        set_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
        desc = np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_per_cam]) if sum(ns) else np.zeros((0, 32), np.uint8)
        o = self.transform_batch(desc, set_off, levelsup)
        bow, node_first, node_id, node_off, idx = [], [0], [], [0], []
        for s in range(len(ns)):
            lo, nw, nf = int(set_off[s]), int(o["n_words"][s]), int(o["n_fv_nodes"][s])
            bow.append((o["bow_ids"][lo:lo + nw].copy(), o["bow_vals"][lo:lo + nw].copy()))
            off = o["fv_off"][lo + s: lo + s + nf + 1]
            node_id.extend(o["fv_node"][lo:lo + nf].tolist())
            base = len(idx)
            idx.extend(o["fv_idx"][lo:lo + int(off[-1]) if nf else lo].tolist())
            node_off.extend((base + off[1:]).tolist())
            node_first.append(len(node_id))
        fv = dict(node_first=np.array(node_first, np.int32), node_id=np.array(node_id, np.int32), node_off=np.array(node_off, np.int32), idx=np.array(idx, np.int32))
        return bow, fv, o
