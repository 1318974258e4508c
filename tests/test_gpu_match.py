"""GPU parity of the Hamming matchers (through the C-ABI) against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import ORBmatcher
import synth

pytestmark = pytest.mark.gpu


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b[:10] = a[:10]
    b[10:20] = ~a[10:20]
    m = ORBmatcher()
    got = m.DescriptorDistance(a, b)
    L = O.lib()
    ref = np.array([L.orc_hamming256(O._ptr(a[i], O.u8p), O._ptr(b[i], O.u8p)) for i in range(len(a))], np.int32)
    assert np.array_equal(got, ref)
    assert got[:10].tolist() == [0] * 10 and got[10:20].tolist() == [256] * 10


@pytest.mark.parametrize("nq,nt", [(1000, 1000), (1, 1), (37, 1500), (1023, 513), (129, 0), (0, 50)])
def test_bruteforce_vs_oracle(nq, nt):
    base = synth.random_descriptors(1, max(nt, 1))
    P, Q, T = 3, max(nq, 1) + 5, max(nt, 1) + 3
    dq = np.zeros((P, Q, 32), np.uint8)
    dt = np.zeros((P, T, 32), np.uint8)
    nqs = np.array([nq, nq, max(nq - 1, 0)], np.int32)
    nts = np.array([nt, max(nt - 1, 0), nt], np.int32)
    rng = np.random.default_rng(nq * 7 + nt)
    for p in range(P):
        dt[p, :nts[p]] = base[:nts[p]]
        if nts[p]:
            src = base[rng.integers(0, nts[p], nqs[p])]
            dq[p, :nqs[p]] = synth.random_descriptors(p, nqs[p], p_flip=0.08, base=src) if nqs[p] else 0
        else:
            dq[p, :nqs[p]] = synth.random_descriptors(p + 9, nqs[p])
    # duplicates in the train set: ties must resolve to the lowest index; second-best equals best
    if nt >= 4:
        dt[0, 3] = dt[0, 1]
    m = ORBmatcher(max_pairs=P, max_query=Q, max_train=T)
    bi, bd, sd = m.bruteforce(dq, nqs, dt, nts)
    for p in range(P):
        rbi, rbd, rsd = O.match_bruteforce(dq[p, :nqs[p]], dt[p, :nts[p]])
        assert np.array_equal(bi[p, :nqs[p]], rbi), (p, nq, nt)
        assert np.array_equal(bd[p, :nqs[p]], rbd)
        assert np.array_equal(sd[p, :nqs[p]], rsd)
        assert (bi[p, nqs[p]:] == -1).all() and (bd[p, nqs[p]:] == 256).all()   # untouched


def test_bruteforce_all_bits_differ():
    """distance 256 never beats the initial bestDist=256 of the reference scan: index stays -1."""
    dq = np.zeros((1, 4, 32), np.uint8)
    dt = np.full((1, 6, 32), 255, np.uint8)
    m = ORBmatcher(max_pairs=1, max_query=4, max_train=6)
    bi, bd, sd = m.bruteforce(dq, np.array([4], np.int32), dt, np.array([6], np.int32))
    rbi, rbd, rsd = O.match_bruteforce(dq[0], dt[0])
    assert np.array_equal(bi[0], rbi) and np.array_equal(bd[0], rbd) and np.array_equal(sd[0], rsd)
    assert (bi == -1).all()


def test_bruteforce_full_config_properties():
    """BASELINE configs[1]: 512 (frame, camera) pairs of 1000 x 1000, device API; properties + sampled oracle parity."""
    import torch
    P, N = 512, 1000
    rng = np.random.default_rng(5)
    dt = rng.integers(0, 256, (P, N, 32), dtype=np.uint8)
    perm = np.stack([rng.permutation(N) for _ in range(P)])
    dq = np.take_along_axis(dt, perm[:, :, None], 1)
    dq ^= np.packbits(rng.random((P, N, 256)) < 0.05, axis=2)
    m = ORBmatcher(max_pairs=P, max_query=N, max_train=N)
    n = torch.full((P,), N, dtype=torch.int32, device="cuda")
    bi, bd, sd = m.bruteforce_device(torch.from_numpy(dq).cuda(), n, torch.from_numpy(dt).cuda(), n)
    torch.cuda.synchronize()
    bi, bd, sd = bi.cpu().numpy(), bd.cpu().numpy(), sd.cpu().numpy()
    assert (bi == perm).mean() > 0.999          # the planted neighbour is found
    assert (bd <= sd).all() and (bd >= 0).all() and (sd <= 256).all()
    for p in (0, 17, 511):
        rbi, rbd, rsd = O.match_bruteforce(dq[p], dt[p])
        assert np.array_equal(bi[p], rbi) and np.array_equal(bd[p], rbd) and np.array_equal(sd[p], rsd)
