"""Host-side mirror of the reference's Optimizer bundle-adjustment entry points (include/Optimizer.h:50-56) above the C-ABI.

    opt = Optimizer()
    poses, points, outlier, stats = opt.LocalBundleAdjustment(problem)            # src/Optimizer.cc:407-696
    poses, points, stats          = opt.GlobalBundleAdjustemnt(problem, nIterations=10, bRobust=True)   # :62-248

`problem` is the flattened graph (dict of numpy arrays, see orbba_problem_t in include/orbslam2_dualcam_b200.h).
"""
import ctypes as C

import numpy as np

from .capi import check, lib, ptr
from .capi import stream_handle as _stream_handle

TH_HUBER_MONO = float(np.float32(np.sqrt(5.991)))    # `const float thHuberMono = sqrt(5.991)`  src/Optimizer.cc:514
CHI2_MONO = 5.991                                     # src/Optimizer.cc:603
TH_HUBER_2D = float(np.float32(np.sqrt(3.99)))        # `const float thHuber2D = sqrt(3.99)`  src/Optimizer.cc:108 (BundleAdjustment / GlobalBundleAdjustemnt)


class ProblemC(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_cams", C.c_int32),
                ("poses", C.c_void_p), ("pose_fixed", C.c_void_p), ("points", C.c_void_p),
                ("edge_pose", C.c_void_p), ("edge_point", C.c_void_p), ("edge_cam", C.c_void_p),
                ("edge_obs", C.c_void_p), ("edge_inv_sigma2", C.c_void_p),
                ("cam_K", C.c_void_p), ("cam_ext", C.c_void_p), ("cam_adj", C.c_void_p)]


class ProblemF32C(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_cams", C.c_int32), ("n_levels", C.c_int32),
                ("poses", C.c_void_p), ("pose_fixed", C.c_void_p), ("points", C.c_void_p), ("edges", C.c_void_p), ("inv_sigma2", C.c_void_p),
                ("cam_K", C.c_void_p), ("cam_ext", C.c_void_p), ("cam_adj", C.c_void_p)]


EDGE16 = np.dtype([("point", "<u4"), ("pose", "<u2"), ("cam", "u1"), ("octave", "u1"), ("u", "<f4"), ("v", "<f4")])


def compact_problem(p, inv_sigma2_levels):
    """dict in the orbba_problem_t layout -> dict in the orbba_problem_f32_t layout (what an adaptor would read straight out of
    KeyFrame / MapPoint): every value of `p` must be float-representable and every weight one of the per-level table."""
    lev = np.asarray(inv_sigma2_levels, np.float32)
    w = np.asarray(p["edge_inv_sigma2"], np.float64)
    octave = np.abs(w[:, None] - lev[None, :].astype(np.float64)).argmin(1)
    assert np.array_equal(lev[octave].astype(np.float64), w), "edge weights are not entries of the level table"
    e = np.zeros(len(w), EDGE16)
    e["point"], e["pose"], e["cam"], e["octave"] = p["edge_point"], p["edge_pose"], p["edge_cam"], octave
    e["u"], e["v"] = p["edge_obs"][:, 0], p["edge_obs"][:, 1]
    for k, a in (("poses", p["poses"]), ("points", p["points"]), ("edge_obs", p["edge_obs"])):
        assert np.array_equal(np.asarray(a, np.float32).astype(np.float64), a), k + " is not float-representable"
    return dict(poses=np.ascontiguousarray(p["poses"], np.float32), pose_fixed=np.ascontiguousarray(p["pose_fixed"], np.uint8),
                points=np.ascontiguousarray(p["points"], np.float32), edges=e, inv_sigma2=lev,
                cam_K=np.ascontiguousarray(p["cam_K"], np.float64), cam_ext=np.ascontiguousarray(p["cam_ext"], np.float64),
                cam_adj=np.ascontiguousarray(p["cam_adj"], np.float64))


class StatsC(C.Structure):
    _fields_ = [("initial_chi2", C.c_double), ("final_chi2", C.c_double), ("final_lambda", C.c_double),
                ("iterations", C.c_int32), ("trials", C.c_int32), ("outliers", C.c_int32), ("status", C.c_int32)]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PoFrameC(C.Structure):
    _fields_ = [("pose", C.c_void_p), ("n_obs", C.c_int32), ("Xw", C.c_void_p), ("obs", C.c_void_p), ("inv_sigma2", C.c_void_p), ("cam", C.c_void_p),
                ("n_cams", C.c_int32), ("cam_K", C.c_void_p), ("cam_ext", C.c_void_p), ("cam_adj", C.c_void_p)]


_KEYS = [("poses", np.float64), ("pose_fixed", np.uint8), ("points", np.float64), ("edge_pose", np.int32), ("edge_point", np.int32),
         ("edge_cam", np.int32), ("edge_obs", np.float64), ("edge_inv_sigma2", np.float64), ("cam_K", np.float64),
         ("cam_ext", np.float64), ("cam_adj", np.float64)]


def problem_struct(p):
    keep = {k: np.ascontiguousarray(p[k], dt) for k, dt in _KEYS}
    s = ProblemC(n_poses=keep["pose_fixed"].shape[0], n_points=keep["points"].shape[0], n_edges=keep["edge_pose"].shape[0],
                 n_cams=keep["cam_K"].shape[0], **{k: keep[k].ctypes.data for k, _ in _KEYS})
    return s, keep


class Optimizer:
    def __init__(self, max_problems=1, device=0):
        self._h = None
        h = C.c_void_p()
        check(lib().orbba_create(C.byref(h), device, max_problems))
        self._h = h
        self._sizes = []

    def close(self):
        if getattr(self, "_h", None):
            lib().orbba_destroy(self._h)
            self._h = None

    __del__ = close

    def LocalBundleAdjustment(self, problem, pbStopFlag=None, its1=5, its2=10, huber_delta=TH_HUBER_MONO, chi2_th=CHI2_MONO):
        """-> (poses [nP][12], points [nL][3], outlier [nE] bool, stats dict); raises OrbError(-5) if the stop flag was set."""
        s, keep = problem_struct(problem)
        poses = np.zeros((s.n_poses, 12)); points = np.zeros((s.n_points, 3)); out = np.zeros(s.n_edges, np.uint8)
        st = StatsC()
        stop = None if pbStopFlag is None else pbStopFlag.ctypes.data
        rc = lib().orbba_local(self._h, C.addressof(s), its1, its2, huber_delta, chi2_th, stop, ptr(poses), ptr(points), ptr(out), C.addressof(st))
        if rc != -5:
            check(rc)
        return poses, points, out.astype(bool), st.asdict()

    def GlobalBundleAdjustemnt(self, problem, nIterations=5, pbStopFlag=None, bRobust=True):
        s, keep = problem_struct(problem)
        poses = np.zeros((s.n_poses, 12)); points = np.zeros((s.n_points, 3))
        st = StatsC()
        stop = None if pbStopFlag is None else pbStopFlag.ctypes.data
        rc = lib().orbba_global(self._h, C.addressof(s), nIterations, TH_HUBER_2D if bRobust else 0.0, stop, ptr(poses), ptr(points), C.addressof(st))
        if rc != -5:
            check(rc)
        return poses, points, st.asdict()

    BundleAdjustment = GlobalBundleAdjustemnt

    def PoseOptimization(self, frames):
        """Optimizer::PoseOptimization(pFrame) (src/Optimizer.cc:250-405) for a batch of frames.
        frames: list of dict(pose [12], Xw [n][3], obs [n][2], inv_sigma2 [n], cam [n], cam_K, cam_ext, cam_adj)
        -> list of (pose [12], outlier bool [n], n_inliers, (lm_iterations, lm_trials))"""
        single = isinstance(frames, dict)
        frames = [frames] if single else list(frames)
        arr = (PoFrameC * len(frames))()
        keeps = []
        for i, f in enumerate(frames):
            k = dict(pose=np.ascontiguousarray(f["pose"], np.float64).reshape(12), Xw=np.ascontiguousarray(f["Xw"], np.float64), obs=np.ascontiguousarray(f["obs"], np.float64),
                     inv_sigma2=np.ascontiguousarray(f["inv_sigma2"], np.float64), cam=np.ascontiguousarray(f["cam"], np.int32),
                     cam_K=np.ascontiguousarray(f["cam_K"], np.float64), cam_ext=np.ascontiguousarray(f["cam_ext"], np.float64), cam_adj=np.ascontiguousarray(f["cam_adj"], np.float64))
            keeps.append(k)
            arr[i] = PoFrameC(pose=k["pose"].ctypes.data, n_obs=len(k["inv_sigma2"]), Xw=k["Xw"].ctypes.data, obs=k["obs"].ctypes.data, inv_sigma2=k["inv_sigma2"].ctypes.data,
                              cam=k["cam"].ctypes.data, n_cams=k["cam_K"].shape[0], cam_K=k["cam_K"].ctypes.data, cam_ext=k["cam_ext"].ctypes.data, cam_adj=k["cam_adj"].ctypes.data)
        n = len(frames)
        tot = sum(a.n_obs for a in arr)
        poses = np.zeros((n, 12)); outl = np.zeros(max(tot, 1), np.uint8); inl = np.zeros(n, np.int32); cnt = np.zeros((n, 2), np.int32)
        check(lib().orbba_pose_optimization(self._h, C.addressof(arr), n, ptr(poses), ptr(outl), ptr(inl), ptr(cnt)))
        res, off = [], 0
        for i, a in enumerate(arr):
            res.append((poses[i].copy(), outl[off:off + a.n_obs].astype(bool), int(inl[i]), (int(cnt[i, 0]), int(cnt[i, 1]))))
            off += a.n_obs
        return res[0] if single else res

    # ---- batched form
    @staticmethod
    def prepare(problems):
        """list of problem dicts -> (ctypes array of orbba_problem_t, keep-alive list); reusable across upload() calls"""
        arr = (ProblemC * len(problems))()
        keeps = []
        for i, p in enumerate(problems):
            s, keep = problem_struct(p)
            arr[i] = s
            keeps.append(keep)
        return arr, keeps

    def upload(self, problems):
        arr, keeps = problems if isinstance(problems, tuple) else self.prepare(problems)
        fn = lib().orbba_upload_f32 if isinstance(arr[0], ProblemF32C) else lib().orbba_upload
        check(fn(self._h, C.addressof(arr), len(arr)))
        self._sizes = [(a.n_poses, a.n_points, a.n_edges) for a in arr]

    @staticmethod
    def prepare_f32(compact_problems):
        """list of compact_problem() dicts -> (ctypes array of orbba_problem_f32_t, keep-alive list) for upload()"""
        arr = (ProblemF32C * len(compact_problems))()
        for i, q in enumerate(compact_problems):
            arr[i] = ProblemF32C(n_poses=len(q["pose_fixed"]), n_points=len(q["points"]), n_edges=len(q["edges"]), n_cams=len(q["cam_K"]),
                                 n_levels=len(q["inv_sigma2"]), **{k: q[k].ctypes.data for k in ("poses", "pose_fixed", "points", "edges", "inv_sigma2", "cam_K", "cam_ext", "cam_adj")})
        return arr, list(compact_problems)

    def set_stream(self, stream):
        h = _stream_handle(stream)
        if getattr(self, "_stream_h", None) != h:
            check(lib().orbba_set_stream(self._h, h))
            self._stream_h = h

    def set_copy_stream(self, stream):
        check(lib().orbba_set_copy_stream(self._h, _stream_handle(stream) if stream is not None else None))

    def run(self, its1=5, its2=10, huber_delta=TH_HUBER_MONO, chi2_th=CHI2_MONO, stream=None):
        if stream is not None:
            self.set_stream(stream)
        check(lib().orbba_run(self._h, its1, its2, huber_delta, chi2_th))

    def download(self, p):
        nP, nL, nE = self._sizes[p]
        poses = np.zeros((nP, 12)); points = np.zeros((nL, 3)); out = np.zeros(nE, np.uint8)
        st = StatsC()
        check(lib().orbba_download(self._h, p, ptr(poses), ptr(points), ptr(out), C.addressof(st)))
        return poses, points, out.astype(bool), st.asdict()

    def download_batch(self, out=None):
        """Results of all problems of the last run, concatenated in upload order -> (poses [sumP][12], points [sumL][3],
        outlier uint8 [sumE], stats list).  `out` = preallocated (poses, points, outlier) numpy arrays / pinned torch tensors."""
        nP = sum(a for a, _, _ in self._sizes); nL = sum(b for _, b, _ in self._sizes); nE = sum(c for _, _, c in self._sizes)
        if out is None:
            out = (np.zeros((nP, 12)), np.zeros((nL, 3)), np.zeros(nE, np.uint8))
        st = (StatsC * len(self._sizes))()
        check(lib().orbba_download_batch(self._h, ptr(out[0]), ptr(out[1]), ptr(out[2]), C.addressof(st)))
        return out[0], out[1], out[2], [s.asdict() for s in st]

    def kernel_ms(self):
        """device ms of {k_lin, k_build, k_land, k_pairs, k_solve, k_back} summed over the recorded LM steps -> (dict, steps)"""
        ms = (C.c_double * 6)()
        n = C.c_int()
        check(lib().orbba_kernel_ms(self._h, ms, C.byref(n)))
        return dict(zip(["k_lin", "k_build", "k_land", "k_pairs", "k_solve", "k_back"], list(ms))), n.value

    def synchronize(self):
        check(lib().orbba_synchronize(self._h))

    def launch_count(self):
        return int(lib().orbba_launch_count(self._h))

    def profile(self, enable=True):
        check(lib().orbba_profile(self._h, int(enable)))

    def stage_ms(self):
        ms, n = C.c_double(), C.c_int()
        check(lib().orbba_stage_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


# ------------------------------------------------------------------------------------------------ distributed global BA
def shard_problem(problem, rank, world):
    """Landmark partition of SURVEY.md §8(e): rank r keeps the map points `index % world == r` and their edges; poses, fixed
    flags and cameras are replicated.  Returns a problem dict whose edge_point indexes the local point array, plus the global
    indices of the kept points / edges (`point_ids`, `edge_ids`) for scattering results back."""
    pts = np.arange(rank, len(problem["points"]), world)
    keep = (problem["edge_point"] % world) == rank
    out = dict(problem)
    out["points"] = np.ascontiguousarray(problem["points"][pts])
    for k in ("edge_pose", "edge_cam", "edge_obs", "edge_inv_sigma2"):
        out[k] = np.ascontiguousarray(problem[k][keep])
    out["edge_point"] = np.ascontiguousarray(problem["edge_point"][keep] // world).astype(np.int32)
    out["point_ids"] = pts
    out["edge_ids"] = np.flatnonzero(keep)
    return out


class DistributedOptimizer:
    """Optimizer::GlobalBundleAdjustemnt over the GPUs of a node (one process per GPU): include/orbslam2_dualcam_b200.h orbba_dist_*."""

    def __init__(self, rank=0, world=1, device=0, unique_id=None):
        self._h = None
        h = C.c_void_p()
        uid = None if unique_id is None else (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(lib().orbba_dist_create(C.byref(h), device, rank, world, uid))
        self._h, self._uid = h, uid
        self.rank, self.world = rank, world

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().orbba_dist_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, device):
        """Creates the library's own NCCL communicator next to an initialised torch.distributed group (the 128-byte id travels
        through broadcast_object_list)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None] if world > 1 else [None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return cls(rank, world, device, box[0])

    def close(self):
        if getattr(self, "_h", None):
            lib().orbba_dist_destroy(self._h)
            self._h = None

    __del__ = close

    def GlobalBundleAdjustemnt(self, shard, nIterations=5, pbStopFlag=None, bRobust=True):
        """collective; -> (poses [nP][12] identical on every rank, local points [nL_local][3], stats)"""
        s, keep = problem_struct(shard)
        poses = np.zeros((s.n_poses, 12)); points = np.zeros((s.n_points, 3))
        st = StatsC()
        stop = None if pbStopFlag is None else pbStopFlag.ctypes.data
        rc = lib().orbba_dist_optimize(self._h, C.addressof(s), nIterations, TH_HUBER_2D if bRobust else 0.0, stop, ptr(poses), ptr(points), C.addressof(st))
        if rc != -5:
            check(rc)
        return poses, points, st.asdict()

    def timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(lib().orbba_dist_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        d, nb = C.c_double(), C.c_longlong()
        check(lib().orbba_dist_loop_ms(self._h, C.byref(d), C.byref(nb)))
        return dict(allreduce_ms=a.value, solve_ms=b.value, allreduce_bytes=c.value, loop_ms=d.value, skyline_blocks=nb.value, segments=int(lib().orbba_dist_segments(self._h)))

    def launch_count(self):
        return int(lib().orbba_dist_launch_count(self._h))
