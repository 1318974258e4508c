"""Pure-python std::list-style restatement of ORBextractor::DistributeOctTree / ExtractorNode::DivideNode
(/root/reference/src/ORBextractor.cc:481-763).  Used by tools/make_golden.py (golden vectors) and by the CPU tests as
an independent second implementation; address tie-break pinned to creation order like the oracle."""
import math

import numpy as np

f32 = np.float32


class Node:
    __slots__ = ("UL", "UR", "BL", "BR", "keys", "nomore", "seq")

    def __init__(self):
        self.keys = []
        self.nomore = False


def divide(n, pts):
    halfX = math.ceil(f32(n.UR[0] - n.UL[0]) / 2)
    halfY = math.ceil(f32(n.BR[1] - n.UL[1]) / 2)
    n1, n2, n3, n4 = Node(), Node(), Node(), Node()
    n1.UL = n.UL; n1.UR = (n.UL[0] + halfX, n.UL[1]); n1.BL = (n.UL[0], n.UL[1] + halfY); n1.BR = (n.UL[0] + halfX, n.UL[1] + halfY)
    n2.UL = n1.UR; n2.UR = n.UR; n2.BL = n1.BR; n2.BR = (n.UR[0], n.UL[1] + halfY)
    n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = n.BL; n3.BR = (n1.BR[0], n.BL[1])
    n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = n.BR
    for k in n.keys:
        x, y = pts[k][0], pts[k][1]
        if x < n1.UR[0]:
            (n1 if y < n1.BR[1] else n3).keys.append(k)
        elif y < n1.BR[1]:
            n2.keys.append(k)
        else:
            n4.keys.append(k)
    for c in (n1, n2, n3, n4):
        c.nomore = len(c.keys) == 1
    return n1, n2, n3, n4


def distribute(pts, minX, maxX, minY, maxY, N):
    """pts: list of (x, y, response) relative coords.  Returns retained indices in list order."""
    nIni = int(math.floor(f32(maxX - minX) / f32(maxY - minY) + 0.5))
    hX = f32(maxX - minX) / f32(nIni)
    nodes = []  # python list used as std::list: index 0 = front
    seq = 0
    ini = []
    for i in range(nIni):
        n = Node()
        n.UL = (int(hX * f32(i)), 0); n.UR = (int(hX * f32(i + 1)), 0)
        n.BL = (n.UL[0], maxY - minY); n.BR = (n.UR[0], maxY - minY)
        n.seq = seq; seq += 1
        nodes.append(n); ini.append(n)
    for k, p in enumerate(pts):
        ini[int(f32(p[0]) / hX)].keys.append(k)
    kept = []
    for n in nodes:
        if len(n.keys) == 1:
            n.nomore = True
        if n.keys:
            kept.append(n)
    nodes = kept
    finish = False
    while not finish:
        prev = len(nodes)
        expandable = []
        n_to_expand = 0
        front = []
        rest = []
        for n in nodes:
            if n.nomore:
                rest.append(n)
                continue
            for c in divide(n, pts):
                if c.keys:
                    c.seq = seq; seq += 1
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        n_to_expand += 1
                        expandable.append(c)
        nodes = front + rest
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + n_to_expand * 3 > N:
            while not finish:
                prev = len(nodes)
                order = sorted(expandable, key=lambda c: (len(c.keys), c.seq))
                expandable = []
                for n in reversed(order):
                    for c in divide(n, pts):
                        if c.keys:
                            c.seq = seq; seq += 1
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                expandable.append(c)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if pts[k][2] > pts[best][2]:
                best = k
        out.append(best)
    return out
