"""GPU parity of bundle adjustment (through the C-ABI) against the FP64 CPU oracle.
Bar (BASELINE.json north_star): every pose within 1e-5 relative (||dT||_F / ||T||_F on the 3x4) and identical outlier sets."""
import numpy as np
import pytest

import oracle_lib as O
from orbslam2_dualcam_b200 import Optimizer, OrbError
import synth

pytestmark = pytest.mark.gpu

POSE_RTOL = 1e-5


def _pose_rel(a, b):
    a, b = a.reshape(-1, 12), b.reshape(-1, 12)
    return (np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max()


def _check(p, got, ref):
    poses, points, out, st = got
    rc, rposes, rpoints, rout, rst = ref
    assert rc == 0
    assert _pose_rel(poses, rposes) <= POSE_RTOL, _pose_rel(poses, rposes)
    assert np.abs(points - rpoints).max() <= 1e-5 * max(1.0, np.abs(rpoints).max())
    assert np.array_equal(out, rout), f"outlier sets differ in {(out != rout).sum()} edges"
    assert st["iterations"] == rst["iterations"] and st["trials"] == rst["trials"]
    assert st["outliers"] == rst["outliers"]
    assert np.isclose(st["initial_chi2"], rst["initial_chi2"], rtol=1e-9)
    assert np.isclose(st["final_chi2"], rst["final_chi2"], rtol=1e-6)


@pytest.mark.parametrize("kw", [dict(seed=1, n_kf=6, n_points=120), dict(seed=2, n_kf=10, n_points=600, n_fixed_extra=3),
                                dict(seed=5, n_kf=3, n_points=40, outlier_frac=0.2), dict(seed=6, n_kf=30, n_points=800),
                                dict(seed=7, n_kf=42, n_points=1000)],
                         ids=["small", "fixed_extra", "many_outliers", "large_window_in_shared_memory", "hs_in_global_memory"])
def test_local_ba_vs_oracle(kw):
    p = synth.ba_problem(**kw)
    opt = Optimizer()
    _check(p, opt.LocalBundleAdjustment(p), O.local_ba(p))


def test_local_ba_config3():
    """BASELINE configs[2]: 20 keyframes x 2 cameras, 4000 points, ~30k edges, 5 + 10 LM iterations."""
    p = synth.ba_problem(0)
    assert 25000 < len(p["edge_pose"]) < 35000
    opt = Optimizer()
    got = opt.LocalBundleAdjustment(p)
    _check(p, got, O.local_ba(p))
    again = opt.LocalBundleAdjustment(p)
    assert np.array_equal(got[0], again[0]) and np.array_equal(got[1], again[1])       # bit-reproducible: no atomics


def test_global_ba_vs_oracle():
    p = synth.ba_problem(4, n_kf=8, n_points=300, outlier_frac=0.02)
    opt = Optimizer()
    poses, points, st = opt.GlobalBundleAdjustemnt(p, nIterations=10, bRobust=True)
    rc, rposes, rpoints, rst = O.global_ba(p, iterations=10)
    assert _pose_rel(poses, rposes) <= POSE_RTOL
    assert st["iterations"] == rst["iterations"] and st["trials"] == rst["trials"]
    poses2, _, st2 = opt.GlobalBundleAdjustemnt(p, nIterations=4, bRobust=False)
    rc, rposes2, _, rst2 = O.global_ba(p, iterations=4, huber_delta=0.0)
    assert _pose_rel(poses2, rposes2) <= POSE_RTOL


def test_stop_flag_set_on_entry():
    p = synth.ba_problem(3, n_kf=5, n_points=100)
    opt = Optimizer()
    stop = np.ones(1, np.uint8)
    poses, points, out, st = opt.LocalBundleAdjustment(p, pbStopFlag=stop)
    assert st["status"] == -5 and st["iterations"] == 0
    assert _pose_rel(poses, p["poses"]) < 1e-6


def test_batched_problems_match_single():
    ps = [synth.ba_problem(10 + i, n_kf=5 + i, n_points=150 + 40 * i) for i in range(5)]
    opt = Optimizer(max_problems=8)
    opt.upload(ps)
    opt.run()
    for i, p in enumerate(ps):
        _check(p, opt.download(i), O.local_ba(p))
    opt.run()          # a second run restarts from the uploaded estimates
    assert np.array_equal(opt.download(2)[0], Optimizer().LocalBundleAdjustment(ps[2])[0])


def test_degenerate_inputs():
    p = synth.ba_problem(8, n_kf=4, n_points=50)
    q = dict(p)
    q["pose_fixed"] = np.ones_like(p["pose_fixed"])          # every pose fixed: structure-only refinement
    opt = Optimizer()
    _check(q, opt.LocalBundleAdjustment(q), O.local_ba(q))
    bad = dict(p)
    bad["edge_pose"] = p["edge_pose"].copy()
    bad["edge_pose"][0] = 99
    with pytest.raises(OrbError):
        opt.LocalBundleAdjustment(bad)


def test_batch_download_copy_stream_and_kernel_timing():
    """orbba_download_batch == per-problem downloads; uploads on a copy stream; per-kernel timing of the LM step."""
    import torch
    ps = [synth.ba_problem(30 + i, n_kf=4 + i, n_points=120 + 30 * i) for i in range(4)]
    opt = Optimizer(max_problems=4)
    copy_stream = torch.cuda.Stream()
    opt.set_copy_stream(copy_stream)
    opt.profile(True)
    opt.upload(opt.prepare(ps))
    opt.run()
    poses, points, outl, stats = opt.download_batch()
    ms, steps = opt.kernel_ms()
    assert steps >= 15 and set(ms) == {"k_lin", "k_build", "k_land", "k_pairs", "k_solve", "k_back"} and all(v >= 0 for v in ms.values())
    oP = oL = oE = 0
    for i, p in enumerate(ps):
        a, b, c, st = opt.download(i)
        nP, nL, nE = len(p["pose_fixed"]), len(p["points"]), len(p["edge_pose"])
        assert np.array_equal(poses[oP:oP + nP], a) and np.array_equal(points[oL:oL + nL], b) and np.array_equal(outl[oE:oE + nE].astype(bool), c)
        assert stats[i]["trials"] == st["trials"]
        _check(p, (a, b, c, st), O.local_ba(p))
        oP += nP; oL += nL; oE += nE
    opt.set_copy_stream(None)


def test_edges_not_grouped_by_landmark():
    """the reference adds edges landmark by landmark; any other order is accepted and the outlier flags come back in caller order"""
    p = synth.ba_problem(12, n_kf=5, n_points=150)
    perm = np.random.default_rng(0).permutation(len(p["edge_pose"]))
    q = dict(p)
    for k in ("edge_pose", "edge_point", "edge_cam", "edge_obs", "edge_inv_sigma2"):
        q[k] = np.ascontiguousarray(p[k][perm])
    a = Optimizer().LocalBundleAdjustment(p)
    b = Optimizer().LocalBundleAdjustment(q)
    assert _pose_rel(b[0], a[0]) <= 1e-9 and np.array_equal(a[2][perm], b[2])      # (the order inside a landmark changes the rounding)


def test_landmark_observed_by_more_key_frames_than_a_block_holds():
    """k_land packs whole landmarks into blocks of 128 edge slots; a landmark with more observations takes the strided path"""
    p = synth.ba_problem(21, n_kf=136, n_points=24, obs_range=(130, 136), outlier_frac=0.02)
    assert np.bincount(p["edge_point"]).max() > 128
    _check(p, Optimizer().LocalBundleAdjustment(p), O.local_ba(p))


def test_heterogeneous_batch_with_rejected_trials():
    """windows of very different sizes, noise levels and outlier shares in one lock-step batch: every window keeps the oracle's own
    trial sequence (rejected trials included) whatever its neighbours do"""
    kws = [dict(seed=40, n_kf=4, n_points=60), dict(seed=41, n_kf=25, n_points=900, outlier_frac=0.15),
           dict(seed=42, n_kf=8, n_points=200, pose_noise=(0.3, 6.0), point_noise=0.5), dict(seed=43, n_kf=12, n_points=400, n_fixed_extra=4),
           dict(seed=44, n_kf=6, n_points=150, pose_noise=(0.5, 10.0), point_noise=1.0, outlier_frac=0.3), dict(seed=45, n_kf=16, n_points=300)]
    ps = [synth.ba_problem(**kw) for kw in kws]
    refs = [O.local_ba(p) for p in ps]
    assert any(r[4]["trials"] > r[4]["iterations"] for r in refs), "the batch should contain rejected LM trials"
    opt = Optimizer(max_problems=len(ps))
    opt.upload(ps)
    opt.run()
    for i, (p, r) in enumerate(zip(ps, refs)):
        _check(p, opt.download(i), r)


def test_compact_f32_upload_equals_f64_upload():
    """orbba_upload_f32 (CV_32F poses / points, 16-byte edge records, per-level weights) gives the bits of orbba_upload"""
    from orbslam2_dualcam_b200 import compact_problem
    ps = [synth.ba_problem(50 + i, n_kf=5 + 2 * i, n_points=150 + 60 * i) for i in range(3)]
    lev = synth.inv_sigma2_levels()
    a = Optimizer(max_problems=3)
    a.upload(ps)
    a.run()
    b = Optimizer(max_problems=3)
    perm = np.random.default_rng(1).permutation(len(ps[1]["edge_pose"]))
    q1 = dict(ps[1])
    for k in ("edge_pose", "edge_point", "edge_cam", "edge_obs", "edge_inv_sigma2"):
        q1[k] = np.ascontiguousarray(ps[1][k][perm])                    # ungrouped edges through the compact path too
    b.upload(Optimizer.prepare_f32([compact_problem(p, lev) for p in (ps[0], q1, ps[2])]))
    b.run()
    for i in (0, 2):
        ra, rb = a.download(i), b.download(i)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and ra[3] == rb[3]
    ra, rb = a.download(1), b.download(1)
    assert _pose_rel(rb[0], ra[0]) <= 1e-9 and np.array_equal(ra[2][perm], rb[2])
