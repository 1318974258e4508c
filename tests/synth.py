"""Seeded synthetic inputs for the hot path (numpy only; shared by tests and bench.py).

Images follow SURVEY.md §8(d): a scene of ~400 random filled rectangles / discs (uniform grey, 8-80 px)
seen through a window that drifts a few pixels per frame, plus per-frame Gaussian noise (sigma 4) and a
sigma-1 blur, so that every pyramid level of a 640x480 frame has a few thousand FAST-20 corners and
consecutive frames share structure (matches exist).  Frame 0 of camera 1 can be replaced by a flat
image to exercise the minThFAST fallback and the empty-level path.
"""
import numpy as np


def make_scene(seed, h, w, n_shapes=400):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n_shapes):
        g = float(rng.integers(0, 256))
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        sx, sy = int(rng.integers(8, 81)), int(rng.integers(8, 81))
        x0, x1 = max(cx - sx // 2, 0), min(cx + sx // 2 + 1, w)
        y0, y1 = max(cy - sy // 2, 0), min(cy + sy // 2 + 1, h)
        if rng.random() < 0.5:
            img[y0:y1, x0:x1] = g
        else:
            r = sx // 2
            m = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = g
    return img


def _blur_sigma1(img):
    k = np.exp(-0.5 * np.arange(-3, 4, dtype=np.float32) ** 2)
    k /= k.sum()
    p = np.pad(img, 3, mode="reflect")
    t = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(7))
    return sum(k[i] * t[i:i + img.shape[0], :] for i in range(7))


def render_frame(scene, ox, oy, W, H, noise_seed):
    rng = np.random.default_rng(noise_seed)
    crop = scene[oy:oy + H, ox:ox + W]
    f = _blur_sigma1(crop + rng.normal(0.0, 4.0, crop.shape).astype(np.float32))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def dual_sequence(seed, n_frames, W=640, H=480, cams=2, flat_first=False, max_shift=8):
    """uint8 [n_frames][cams][H][W]: one drifting window per camera over its own scene."""
    out = np.empty((n_frames, cams, H, W), np.uint8)
    margin = 64
    for c in range(cams):
        scene = make_scene(seed * 131 + c, H + 2 * margin, W + 2 * margin)
        rng = np.random.default_rng(seed * 977 + c + 17)
        ox, oy = margin, margin
        for k in range(n_frames):
            out[k, c] = render_frame(scene, ox, oy, W, H, (seed * 1000003 + k) * 2 + c)
            ox = int(np.clip(ox + rng.integers(-max_shift, max_shift + 1), 0, 2 * margin))
            oy = int(np.clip(oy + rng.integers(-max_shift, max_shift + 1), 0, 2 * margin))
    if flat_first and cams > 1:
        out[0, 1] = 97
    return out


def tiled_batch(seed, n_frames, W=640, H=480, cams=2, unique=16):
    """Cheap large batch for benchmarking: `unique` rendered dual-frames, the rest are cyclic shifts of them
    (every frame distinct as a byte string, same corner statistics)."""
    base = dual_sequence(seed, min(unique, n_frames), W, H, cams)
    out = np.empty((n_frames, cams, H, W), np.uint8)
    for k in range(n_frames):
        b = base[k % base.shape[0]]
        r = k // base.shape[0]
        out[k] = np.roll(b, shift=(3 * r, 5 * r), axis=(1, 2)) if r else b
    return out


def random_descriptors(seed, n, p_flip=None, base=None):
    """n x 32 uint8 descriptors; with `base` given, a noisy copy (each bit flipped with prob p_flip)."""
    rng = np.random.default_rng(seed)
    if base is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    flips = np.packbits(rng.random((base.shape[0], 256)) < p_flip, axis=1)
    return base ^ flips


# ------------------------------------------------------------------------------------------------ bundle adjustment
# Dual rig of the reference's example settings (Dual-LenaCV.yaml:12-44): intrinsics of both cameras, extrinsic of camera 1.
RIG_K = np.array([[558.4684, 560.0944, 326.7993, 262.9017],
                  [546.597961663159, 546.254254416417, 332.758939924785, 247.385425357685]], np.float64)
RIG_Q1 = (0.82351, -0.00262741, 0.567257, 0.00665084)     # qw qx qy qz
RIG_T1 = (0.069481, -0.000909887, -0.0713882)


def _quat_to_R(qw, qx, qy, qz):
    n = np.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    qw, qx, qy, qz = qw / n, qx / n, qy / n, qz / n
    return np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                     [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                     [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def rig_extrinsics():
    """(ext [2][3][4], adj [2][6][6]) as the reference holds them: float32 matrices; the 6x6 'adjoint' of
    src/Cameras.cc:26-40 (upper-left R, lower-right R, upper-right R*[t]x; the lower-left block, which the reference never
    writes, is pinned to 0)."""
    ext = np.zeros((2, 3, 4), np.float32)
    ext[0, :, :3] = np.eye(3)
    ext[1, :, :3] = _quat_to_R(*RIG_Q1)
    ext[1, :, 3] = RIG_T1
    adj = np.zeros((2, 6, 6), np.float32)
    for c in range(2):
        R, t = ext[c, :, :3], ext[c, :, 3]
        t_hat = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)
        adj[c, :3, :3] = R
        adj[c, 3:, 3:] = R
        adj[c, :3, 3:] = (R @ t_hat).astype(np.float32)
    return ext.astype(np.float64), adj.astype(np.float64)


def inv_sigma2_levels(n_levels=8, f=1.2):
    """ORBextractor::mvInvLevelSigma2 (src/ORBextractor.cc:419-431): float products, as ba_problem weights its observations"""
    out = np.ones(n_levels, np.float32)
    sc = np.float32(1.0)
    for l in range(1, n_levels):
        sc = np.float32(sc * np.float32(f))
        out[l] = np.float32(1.0) / np.float32(sc * sc)
    return out


def ba_problem(seed=0, n_kf=20, n_points=4000, n_fixed_extra=0, obs_range=(5, 10), outlier_frac=0.05, W=640, H=480,
               pose_noise=(0.02, 0.5), point_noise=0.05):
    """Synthetic LocalBundleAdjustment input.  Returns a dict of numpy arrays in the layout of
    orbba_problem_t plus the ground truth.  Pose 0 is the fixed one (fixId); `n_fixed_extra` more poses at the end are
    fixed too (the reference's lFixedCameras)."""
    rng = np.random.default_rng(seed)
    ext, adj = rig_extrinsics()
    nP = n_kf + n_fixed_extra
    # rig trajectory (world <- rig), smooth: forward motion with gentle yaw/pitch
    Twc = []
    for i in range(nP):
        s = i / max(nP - 1, 1)
        R = _rodrigues(np.array([0.03 * np.sin(2 * s), 0.25 * s - 0.1, 0.02 * s]))
        t = np.array([1.5 * s, 0.1 * np.sin(3 * s), 2.0 * s])
        Twc.append((R, t))
    gt_poses = np.zeros((nP, 3, 4))
    for i, (R, t) in enumerate(Twc):
        gt_poses[i, :, :3] = R.T
        gt_poses[i, :, 3] = -R.T @ t

    def project(Tcw, c, X):
        pr = Tcw[:, :3] @ X + Tcw[:, 3]
        pc = ext[c, :, :3] @ pr + ext[c, :, 3]
        if pc[2] <= 0.1:
            return None
        u = RIG_K[c, 0] * pc[0] / pc[2] + RIG_K[c, 2]
        v = RIG_K[c, 1] * pc[1] / pc[2] + RIG_K[c, 3]
        if 0 <= u < W and 0 <= v < H:
            return np.array([u, v])
        return None

    inv_sigma2 = np.ones(8, np.float32)
    sc = np.float32(1.0)
    for l in range(1, 8):
        sc = np.float32(sc * np.float32(1.2))
        inv_sigma2[l] = np.float32(1.0) / np.float32(sc * sc)
    pts, e_pose, e_pt, e_cam, e_obs, e_info, e_out = [], [], [], [], [], [], []
    while len(pts) < n_points:
        k, c = int(rng.integers(0, nP)), int(rng.integers(0, 2))
        u, v, d = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(2, 20)
        pc = np.array([(u - RIG_K[c, 2]) / RIG_K[c, 0] * d, (v - RIG_K[c, 3]) / RIG_K[c, 1] * d, d])
        pr = ext[c, :, :3].T @ (pc - ext[c, :, 3])
        X = gt_poses[k, :, :3].T @ (pr - gt_poses[k, :, 3])
        vis = []
        for kk in rng.permutation(nP):
            for cc in rng.permutation(2):
                p = project(gt_poses[kk], cc, X)
                if p is not None:
                    vis.append((int(kk), int(cc), p))
                    break      # one observation per keyframe (MapPoint::mObservations is a map<KeyFrame, idx>)
        want = int(rng.integers(obs_range[0], obs_range[1] + 1))
        if len(vis) < 2:
            continue
        vis = vis[:want]
        pid = len(pts)
        pts.append(X)
        for kk, cc, p in vis:
            octave = int(rng.integers(0, 8))
            noise = rng.normal(0, 1.2 ** octave, 2)
            out = rng.random() < outlier_frac
            if out:
                noise = noise + rng.choice([-20.0, 20.0], 2)
            e_pose.append(kk); e_pt.append(pid); e_cam.append(cc)
            e_obs.append((p + noise).astype(np.float32))   # cv::KeyPoint::pt is float
            e_info.append(inv_sigma2[octave]); e_out.append(out)
    gt_points = np.array(pts)
    # initial estimates: perturbed, then rounded to float32 (the reference keeps poses / points in CV_32F)
    poses = gt_poses.copy()
    for i in range(1, n_kf):
        dR = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]) / np.sqrt(3), 3))
        poses[i, :, :3] = dR @ poses[i, :, :3]
        poses[i, :, 3] = dR @ poses[i, :, 3] + rng.normal(0, pose_noise[0] / np.sqrt(3), 3)
    points = gt_points + rng.normal(0, point_noise / np.sqrt(3), gt_points.shape)
    fixed = np.zeros(nP, np.uint8)
    fixed[0] = 1
    fixed[n_kf:] = 1
    order = np.lexsort((np.array(e_pose), np.array(e_pt)))   # edges listed per point, as the reference adds them
    return dict(
        poses=np.ascontiguousarray(poses.astype(np.float32).astype(np.float64).reshape(nP, 12)), pose_fixed=fixed,
        points=np.ascontiguousarray(points.astype(np.float32).astype(np.float64)),
        edge_pose=np.array(e_pose, np.int32)[order], edge_point=np.array(e_pt, np.int32)[order], edge_cam=np.array(e_cam, np.int32)[order],
        edge_obs=np.ascontiguousarray(np.array(e_obs, np.float64)[order]), edge_inv_sigma2=np.array(e_info, np.float64)[order],
        cam_K=RIG_K.astype(np.float32).astype(np.float64), cam_ext=np.ascontiguousarray(ext.reshape(2, 12)), cam_adj=np.ascontiguousarray(adj.reshape(2, 36)),
        gt_poses=gt_poses.reshape(nP, 12), gt_points=gt_points, planted_outlier=np.array(e_out, bool)[order])


# ------------------------------------------------------------------------------------------------ guided searches
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
MP_DTYPE = np.dtype([("valid", "<i4"), ("cam", "<i4"), ("u", "<f4"), ("v", "<f4"), ("level", "<i4"), ("view_cos", "<f4"),
                     ("obs_positive", "<i4"), ("desc", "u1", (32,))])


def scale_factors(n_levels=8, f=1.2):
    s = np.ones(n_levels, np.float32)
    for i in range(1, n_levels):
        s[i] = np.float32(s[i - 1] * np.float32(f))
    return s


def search_frame(seed, n_kp=(1000, 1000), W=640, H=480, n_levels=8, clustered=True):
    """A dual Frame as the searches see it: undistorted keypoints (a few fall outside the image bounds, as undistortion can
    produce), descriptors, bounds, scale factors."""
    rng = np.random.default_rng(seed)
    total = int(sum(n_kp))
    kps = np.zeros(total, KP_DTYPE)
    off = 0
    for n in n_kp:
        if clustered and n > 0:
            centres = rng.uniform([20, 20], [W - 20, H - 20], (max(n // 25, 1), 2))
            pts = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 12, (n, 2))
        else:
            pts = rng.uniform([-3, -3], [W + 3, H + 3], (n, 2))
        kps["x"][off:off + n] = pts[:, 0]
        kps["y"][off:off + n] = pts[:, 1]
        off += n
    kps["octave"] = rng.integers(0, n_levels, total)
    kps["angle"] = rng.uniform(0, 360, total)
    kps["size"] = 31
    kps["class_id"] = -1
    bounds = np.tile(np.array([0, W, 0, H], np.float32), (len(n_kp), 1))
    return dict(n_kp=np.array(n_kp, np.int32), kps_un=kps, desc=random_descriptors(seed + 1, total), bounds=bounds, scale_factors=scale_factors(n_levels))


def local_map_points(seed, frame, n_mp, p_flip=0.08, frac_invalid=0.1, frac_obs0=0.05):
    """Map points of TrackLocalMap projected into `frame`: most sit near an existing keypoint with a noisy copy of its
    descriptor (so that windows hold the true match plus distractors), some project to empty areas."""
    rng = np.random.default_rng(seed)
    kps, n_kp = frame["kps_un"], frame["n_kp"]
    first = np.concatenate([[0], np.cumsum(n_kp)])
    total = int(first[-1])
    mps = np.zeros(n_mp, MP_DTYPE)
    if total == 0:
        mps["valid"] = 1
        mps["u"] = rng.uniform(0, 640, n_mp); mps["v"] = rng.uniform(0, 480, n_mp)
        mps["level"] = rng.integers(0, len(frame["scale_factors"]), n_mp)
        mps["desc"] = random_descriptors(seed + 3, n_mp)
        return mps
    g = rng.integers(0, total, n_mp)
    cam = np.searchsorted(first, g, side="right") - 1
    mps["cam"] = cam
    mps["u"] = kps["x"][g] + rng.normal(0, 2.5, n_mp)
    mps["v"] = kps["y"][g] + rng.normal(0, 2.5, n_mp)
    nl = len(frame["scale_factors"])
    mps["level"] = np.clip(kps["octave"][g] + rng.integers(-1, 2, n_mp), 0, nl - 1)
    mps["view_cos"] = rng.choice(np.array([0.9, 0.9985, 0.999, 0.7], np.float32), n_mp)
    mps["desc"] = random_descriptors(seed + 2, n_mp, p_flip, frame["desc"][g])
    far = rng.random(n_mp) < 0.1
    mps["u"][far] = rng.uniform(-50, 700, far.sum()); mps["v"][far] = rng.uniform(-50, 530, far.sum())
    mps["valid"] = rng.random(n_mp) >= frac_invalid
    mps["obs_positive"] = rng.random(n_mp) >= frac_obs0
    return mps


def motion_model_scene(seed, n_pts=(900, 800), W=640, H=480, n_levels=8, rot_outliers=0.15, p_flip=0.06):
    """TrackWithMotionModel inputs: 3-D points seen by each camera of the rig in the last frame, the current frame's keypoints
    (their projections under the current pose + noise + distractors) and the per-camera Rsw / tsw / K of the current frame."""
    rng = np.random.default_rng(seed)
    ext, _ = rig_extrinsics()
    Rcw = _rodrigues(np.array([0.01, -0.02, 0.005]))
    tcw = np.array([0.03, -0.01, 0.05])
    C = len(n_pts)
    Rsw = np.zeros((C, 9), np.float32); tsw = np.zeros((C, 3), np.float32)
    K = RIG_K[:C].astype(np.float32)
    cur_kps, cur_desc, last = [], [], dict(cam=[], valid=[], pos=[], desc=[], octave=[], angle=[], obs_positive=[])
    for c in range(C):
        R = (ext[c, :, :3] @ Rcw).astype(np.float32); t = (ext[c, :, :3] @ tcw + ext[c, :, 3]).astype(np.float32)
        Rsw[c] = R.reshape(9); tsw[c] = t
        n = n_pts[c]
        uv = rng.uniform([5, 5], [W - 5, H - 5], (n, 2)); d = rng.uniform(2, 15, n)
        pc = np.stack([(uv[:, 0] - K[c, 2]) / K[c, 0] * d, (uv[:, 1] - K[c, 3]) / K[c, 1] * d, d], 1)
        Xw = (pc - t) @ R.astype(np.float64)          # R^T (pc - t)
        k = np.zeros(n + n // 3, KP_DTYPE)
        k["x"][:n] = uv[:, 0] + rng.normal(0, 1.5, n); k["y"][:n] = uv[:, 1] + rng.normal(0, 1.5, n)
        k["x"][n:] = rng.uniform(0, W, n // 3); k["y"][n:] = rng.uniform(0, H, n // 3)
        k["octave"] = rng.integers(0, n_levels, len(k)); k["angle"] = rng.uniform(0, 360, len(k)); k["class_id"] = -1
        dsc = random_descriptors(seed * 7 + c, len(k))
        cur_kps.append(k); cur_desc.append(dsc)
        ang = (k["angle"][:n] + 12.0 + rng.normal(0, 2, n)) % 360
        bad = rng.random(n) < rot_outliers
        ang[bad] = rng.uniform(0, 360, bad.sum())
        last["cam"].append(np.full(n, c)); last["valid"].append(rng.random(n) > 0.1); last["pos"].append(Xw)
        last["desc"].append(random_descriptors(seed * 11 + c, n, p_flip, dsc[:n]))
        last["octave"].append(np.clip(k["octave"][:n] + rng.integers(-1, 2, n), 0, n_levels - 1)); last["angle"].append(ang)
        last["obs_positive"].append(rng.random(n) > 0.03)
    cur = dict(n_kp=np.array([len(k) for k in cur_kps], np.int32), kps_un=np.concatenate(cur_kps), desc=np.concatenate(cur_desc),
               bounds=np.tile(np.array([0, W, 0, H], np.float32), (C, 1)), scale_factors=scale_factors(n_levels))
    lastd = {k: np.concatenate(v) for k, v in last.items()}
    perm = rng.permutation(len(lastd["cam"]))         # last-frame keypoints of the two cameras interleave in no particular order here
    lastd = {k: v[np.sort(perm)] if k == "_" else v for k, v in lastd.items()}
    return cur, Rsw, tsw, K, lastd


def bow_scene(seed, n_kp=(1000, 900), n_nodes=90, p_flip=0.05, frac_valid=0.6):
    """SearchByBoW inputs: a KeyFrame and a Frame whose features are noisy copies of key-frame features that mostly fall into the
    same vocabulary node (DBoW2::FeatureVector flattened to CSR: node ids ascending, camera-local indices ascending)."""
    rng = np.random.default_rng(seed)
    C = len(n_kp)

    def csr(nodes_per_cam):
        node_first, node_id, node_off, idx = [0], [], [0], []
        for nodes in nodes_per_cam:
            for nid in np.unique(nodes):
                node_id.append(int(nid)); idx.extend(np.flatnonzero(nodes == nid).tolist()); node_off.append(len(idx))
            node_first.append(len(node_id))
        return dict(node_first=np.array(node_first, np.int32), node_id=np.array(node_id, np.int32), node_off=np.array(node_off, np.int32),
                    idx=np.array(idx, np.int32))

    kdesc, knodes, fdesc, fnodes, kang, fang = [], [], [], [], [], []
    for c, n in enumerate(n_kp):
        d = random_descriptors(seed * 31 + c, n)
        nodes = rng.integers(0, n_nodes, n) * 7 + 3           # sparse node ids
        nf = int(n * 1.1)
        pick = rng.integers(0, max(n, 1), nf)
        fd = random_descriptors(seed * 37 + c, nf, p_flip, d[pick]) if n else random_descriptors(seed * 37 + c, nf)
        fn = nodes[pick].copy() if n else rng.integers(0, n_nodes, nf) * 7 + 3
        stray = rng.random(nf) < 0.1
        fn[stray] = rng.integers(0, n_nodes + 5, stray.sum()) * 7 + 3
        ka = rng.normal(75.0, 3, n) % 360
        bad = rng.random(n) < 0.1
        ka[bad] = rng.uniform(0, 360, bad.sum())
        kdesc.append(d); knodes.append(nodes); fdesc.append(fd); fnodes.append(fn); kang.append(ka); fang.append(rng.normal(30.0 + 100.0 * c, 3, nf) % 360)
    KF = dict(n_kp=np.array(n_kp, np.int32), desc=np.concatenate(kdesc), angle=np.concatenate(kang).astype(np.float32), **csr(knodes))
    F = dict(n_kp=np.array([len(x) for x in fdesc], np.int32), desc=np.concatenate(fdesc), angle=np.concatenate(fang).astype(np.float32), **csr(fnodes))
    valid = (rng.random(int(sum(n_kp))) < frac_valid).astype(np.uint8)
    return F, KF, valid


def frustum_scene(seed, n=3000, W=640, H=480, n_levels=8):
    rng = np.random.default_rng(seed)
    ext, _ = rig_extrinsics()
    Rcw = _rodrigues(np.array([0.02, 0.05, -0.01])); tcw = np.array([0.1, 0.0, -0.2])
    Rsw = np.zeros((2, 9), np.float32); tsw = np.zeros((2, 3), np.float32); Ow = np.zeros((2, 3), np.float32)
    for c in range(2):
        R = ext[c, :, :3] @ Rcw; t = ext[c, :, :3] @ tcw + ext[c, :, 3]
        Rsw[c] = R.reshape(9); tsw[c] = t; Ow[c] = -R.T @ t
    pos = rng.uniform([-15, -8, -5], [15, 8, 25], (n, 3)).astype(np.float32)
    normal = pos - Ow[0] + rng.normal(0, 3.0, (n, 3))
    normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
    d = np.linalg.norm(pos - Ow[0], axis=1)
    max_dist = (d * rng.uniform(0.6, 3.0, n)).astype(np.float32)
    min_dist = (max_dist / 1.2 ** (n_levels - 1)).astype(np.float32)
    frame = dict(Rsw=Rsw, tsw=tsw, Ow=Ow, K=RIG_K.astype(np.float32), bounds=np.tile(np.array([0, W, 0, H], np.float32), (2, 1)), n_levels=n_levels,
                 log_scale_factor=np.float32(np.log(np.float32(1.2))))
    return frame, pos, normal, max_dist, min_dist


def gba_problem(seed=0, n_kf=2000, n_points=200000, obs=8, outlier_frac=0.02, W=640, H=480, pose_noise=(0.01, 0.2), point_noise=0.03):
    """Synthetic GlobalBundleAdjustemnt input, vectorised: a long smooth rig trajectory, every map point
    placed in front of a random key frame and observed by up to `obs` neighbouring key frames (whichever camera of the rig sees it).
    Same dict layout as ba_problem; pose 0 is fixed."""
    rng = np.random.default_rng(seed)
    ext, adj = rig_extrinsics()
    s = np.arange(n_kf) / max(n_kf - 1, 1)
    L = 0.15 * n_kf                                            # 15 cm between key frames
    gt = np.zeros((n_kf, 3, 4))
    for i in range(n_kf):
        R = _rodrigues(np.array([0.02 * np.sin(6 * s[i]), 0.8 * np.sin(3 * s[i]), 0.01 * s[i]]))
        t = np.array([L * s[i], 0.2 * np.sin(5 * s[i]), 0.3 * L * np.sin(2 * s[i])])
        gt[i, :, :3] = R.T
        gt[i, :, 3] = -R.T @ t
    K = RIG_K
    k0 = rng.integers(0, n_kf, n_points); c0 = rng.integers(0, 2, n_points)
    u = rng.uniform(40, W - 40, n_points); v = rng.uniform(40, H - 40, n_points); d = rng.uniform(2, 12, n_points)
    pc = np.stack([(u - K[c0, 2]) / K[c0, 0] * d, (v - K[c0, 3]) / K[c0, 1] * d, d], 1)
    pr = np.einsum("nji,nj->ni", ext[c0][:, :, :3], pc - ext[c0][:, :, 3])
    X = np.einsum("nji,nj->ni", gt[k0][:, :, :3], pr - gt[k0][:, :, 3])
    inv_sigma2 = (1.0 / (scale_factors(8).astype(np.float64) ** 2)).astype(np.float32)
    e_pose, e_pt, e_cam, e_obs = [], [], [], []
    for dk in range(-(obs // 2), obs - obs // 2):
        kk = k0 + dk
        ok = (kk >= 0) & (kk < n_kf)
        kk = np.clip(kk, 0, n_kf - 1)
        prr = np.einsum("nij,nj->ni", gt[kk][:, :, :3], X) + gt[kk][:, :, 3]
        seen = np.zeros(n_points, bool)
        for c in (0, 1):
            pcc = prr @ ext[c, :, :3].T + ext[c, :, 3]
            z = np.where(pcc[:, 2] > 0.2, pcc[:, 2], 1.0)
            uu = K[c, 0] * pcc[:, 0] / z + K[c, 2]; vv = K[c, 1] * pcc[:, 1] / z + K[c, 3]
            vis = ok & ~seen & (pcc[:, 2] > 0.2) & (uu >= 0) & (uu < W) & (vv >= 0) & (vv < H)
            idx = np.flatnonzero(vis)
            e_pose.append(kk[idx]); e_pt.append(idx); e_cam.append(np.full(len(idx), c)); e_obs.append(np.stack([uu[idx], vv[idx]], 1))
            seen |= vis
    e_pose = np.concatenate(e_pose); e_pt = np.concatenate(e_pt); e_cam = np.concatenate(e_cam); e_obs = np.concatenate(e_obs)
    octave = rng.integers(0, 8, len(e_pose))
    noise = rng.normal(0, 1, (len(e_pose), 2)) * (1.2 ** octave)[:, None]
    out = rng.random(len(e_pose)) < outlier_frac
    noise[out] += rng.choice([-20.0, 20.0], (int(out.sum()), 2))
    e_obs = (e_obs + noise).astype(np.float32).astype(np.float64)
    order = np.lexsort((e_pose, e_pt))
    poses = gt.copy()
    for i in range(1, n_kf):
        dR = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]) / np.sqrt(3), 3))
        poses[i, :, :3] = dR @ poses[i, :, :3]
        poses[i, :, 3] = dR @ poses[i, :, 3] + rng.normal(0, pose_noise[0] / np.sqrt(3), 3)
    points = X + rng.normal(0, point_noise / np.sqrt(3), X.shape)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1
    return dict(poses=np.ascontiguousarray(poses.astype(np.float32).astype(np.float64).reshape(n_kf, 12)), pose_fixed=fixed,
                points=np.ascontiguousarray(points.astype(np.float32).astype(np.float64)),
                edge_pose=e_pose[order].astype(np.int32), edge_point=e_pt[order].astype(np.int32), edge_cam=e_cam[order].astype(np.int32),
                edge_obs=np.ascontiguousarray(e_obs[order]), edge_inv_sigma2=inv_sigma2[octave][order].astype(np.float64),
                cam_K=RIG_K.astype(np.float32).astype(np.float64), cam_ext=np.ascontiguousarray(ext.reshape(2, 12)), cam_adj=np.ascontiguousarray(adj.reshape(2, 36)),
                gt_poses=gt.reshape(n_kf, 12), gt_points=X, planted_outlier=out[order])


def pose_opt_frame(seed, n_obs=600, outlier_frac=0.15, W=640, H=480, pose_noise=(0.05, 1.5)):
    """Optimizer::PoseOptimization input: a rig pose perturbed from the truth, map points seen by either camera of the rig,
    octave-dependent pixel noise and gross outliers.  dict in the layout of orbpo_frame_t (+ gt_pose, planted)."""
    rng = np.random.default_rng(seed)
    ext, adj = rig_extrinsics()
    R = _rodrigues(np.array([0.1, -0.2, 0.05])); t = np.array([0.3, -0.1, 0.5])
    gt = np.concatenate([R, t[:, None]], 1)
    cam = rng.integers(0, 2, n_obs)
    u = rng.uniform(10, W - 10, n_obs); v = rng.uniform(10, H - 10, n_obs); d = rng.uniform(1.5, 15, n_obs)
    K = RIG_K
    pc = np.stack([(u - K[cam, 2]) / K[cam, 0] * d, (v - K[cam, 3]) / K[cam, 1] * d, d], 1)
    pr = np.einsum("nji,nj->ni", ext[cam][:, :, :3], pc - ext[cam][:, :, 3])
    Xw = (pr - t) @ R
    octave = rng.integers(0, 8, n_obs)
    noise = rng.normal(0, 1, (n_obs, 2)) * (1.2 ** octave)[:, None]
    planted = rng.random(n_obs) < outlier_frac
    noise[planted] += rng.choice([-25.0, 25.0], (int(planted.sum()), 2))
    obs = (np.stack([u, v], 1) + noise).astype(np.float32).astype(np.float64)
    dR = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]) / np.sqrt(3), 3))
    pose = np.concatenate([dR @ R, (dR @ t + rng.normal(0, pose_noise[0] / np.sqrt(3), 3))[:, None]], 1)
    inv_sigma2 = (1.0 / (scale_factors(8).astype(np.float64) ** 2)).astype(np.float32).astype(np.float64)[octave]
    return dict(pose=pose.astype(np.float32).astype(np.float64).reshape(12), Xw=Xw.astype(np.float32).astype(np.float64), obs=obs, inv_sigma2=inv_sigma2,
                cam=cam.astype(np.int32), cam_K=RIG_K.astype(np.float32).astype(np.float64), cam_ext=np.ascontiguousarray(ext.reshape(2, 12)),
                cam_adj=np.ascontiguousarray(adj.reshape(2, 36)), gt_pose=gt.reshape(12), planted=planted)


# ------------------------------------------------------------------------------------------------ key-frame flavoured searches
def _csr(nodes_per_cam):
    """per-camera node id of every feature -> DBoW2::FeatureVector flattened to CSR (node ids ascending, camera-local indices ascending)"""
    node_first, node_id, node_off, idx = [0], [], [0], []
    for nodes in nodes_per_cam:
        for nid in np.unique(nodes):
            node_id.append(int(nid)); idx.extend(np.flatnonzero(nodes == nid).tolist()); node_off.append(len(idx))
        node_first.append(len(node_id))
    return dict(node_first=np.array(node_first, np.int32), node_id=np.array(node_id, np.int32), node_off=np.array(node_off, np.int32),
                idx=np.array(idx, np.int32))


def _rig_view(Rcw, tcw, W, H, n_levels):
    ext, _ = rig_extrinsics()
    Rsw = np.zeros((2, 9), np.float32); tsw = np.zeros((2, 3), np.float32); Ow = np.zeros((2, 3), np.float32)
    for c in range(2):
        R = ext[c, :, :3] @ Rcw; t = ext[c, :, :3] @ tcw + ext[c, :, 3]
        Rsw[c] = R.reshape(9); tsw[c] = t; Ow[c] = -R.T @ t
    return dict(Rsw=Rsw, tsw=tsw, Ow=Ow, K=RIG_K.astype(np.float32), bounds=np.tile(np.array([0, W, 0, H], np.float32), (2, 1)), n_levels=n_levels,
                log_scale_factor=np.float32(np.log(np.float32(1.2))))


def kf_projection_scene(seed, n_kp=(1000, 900), n_stray=300, W=640, H=480, n_levels=8, p_flip=0.06, pix_noise=1.5, frac_invalid=0.1):
    """A dual key frame (or frame), its per-camera view and candidate map points for the projection searches of relocalisation /
    loop closing / fusion: most points are back-projections of key points (noisy descriptor copy, scale range consistent with the
    key point's octave, normal along the viewing ray), some are strays anywhere around the rig.
    -> (frame, view, points, blocked uint8 [totalN])"""
    rng = np.random.default_rng(seed)
    frame = search_frame(seed, n_kp, W, H, n_levels, clustered=False)
    view = _rig_view(_rodrigues(np.array([0.015, -0.03, 0.01])), np.array([0.05, -0.02, 0.08]), W, H, n_levels)
    first = np.concatenate([[0], np.cumsum(n_kp)])
    total = int(first[-1])
    kps = frame["kps_un"]
    pos, normal, maxd, mind, desc, ang = [], [], [], [], [], []
    for c in range(len(n_kp)):
        R = view["Rsw"][c].reshape(3, 3).astype(np.float64); t = view["tsw"][c].astype(np.float64); K = view["K"][c].astype(np.float64)
        pick = rng.permutation(n_kp[c])[: int(n_kp[c] * 0.7)] + first[c]
        n = len(pick)
        d = rng.uniform(2, 20, n)
        u = kps["x"][pick] + rng.normal(0, pix_noise, n); v = kps["y"][pick] + rng.normal(0, pix_noise, n)
        pc = np.stack([(u - K[2]) / K[0] * d, (v - K[3]) / K[1] * d, d], 1)
        Xw = (pc - t) @ R
        O = view["Ow"][c].astype(np.float64)
        dist = np.linalg.norm(Xw - O, axis=1)
        lvl = kps["octave"][pick] + rng.choice([-1, 0, 0, 0, 1], n) - rng.uniform(0.1, 0.9, n)
        pos.append(Xw); maxd.append(dist * 1.2 ** lvl); mind.append(dist * 1.2 ** lvl / 1.2 ** (n_levels - 1))
        nr = (Xw - O) / dist[:, None] + rng.normal(0, 0.5, (n, 3))
        normal.append(nr / np.linalg.norm(nr, axis=1, keepdims=True))
        desc.append(random_descriptors(seed * 13 + c, n, p_flip, frame["desc"][pick]))
        a = (kps["angle"][pick] + 20.0 + rng.normal(0, 2, n)) % 360
        bad = rng.random(n) < 0.12
        a[bad] = rng.uniform(0, 360, bad.sum())
        ang.append(a)
    Xs = rng.uniform([-20, -10, -10], [20, 10, 30], (n_stray, 3))
    ds = np.linalg.norm(Xs - view["Ow"][0], axis=1)
    pos.append(Xs); maxd.append(ds * rng.uniform(0.7, 3.0, n_stray)); mind.append(maxd[-1] / 1.2 ** (n_levels - 1))
    nr = rng.normal(0, 1, (n_stray, 3)); normal.append(nr / np.linalg.norm(nr, axis=1, keepdims=True))
    desc.append(random_descriptors(seed * 17, n_stray)); ang.append(rng.uniform(0, 360, n_stray))
    N = sum(len(p) for p in pos)
    perm = rng.permutation(N)
    points = dict(valid=(rng.random(N) >= frac_invalid).astype(np.uint8), pos=np.concatenate(pos)[perm].astype(np.float32),
                  normal=np.concatenate(normal)[perm].astype(np.float32), max_dist=np.concatenate(maxd)[perm].astype(np.float32),
                  min_dist=np.concatenate(mind)[perm].astype(np.float32), desc=np.ascontiguousarray(np.concatenate(desc)[perm]),
                  angle=np.concatenate(ang)[perm].astype(np.float32))
    blocked = (rng.random(total) < 0.15).astype(np.uint8)
    return frame, view, points, blocked


def triangulation_scene(seed, n=900, cam=0, W=640, H=480, n_levels=8, n_nodes=80, p_flip=0.05):
    """SearchForTriangulation inputs: two dual key frames seeing the same 3-D points from two poses, key points = projections + noise,
    shared vocabulary nodes for true correspondences, the fundamental matrix F12 of camera `cam` (LocalMapping::ComputeF12 layout:
    x1^T F12 x2 = 0) and the quantities of the epipole.  -> dict"""
    rng = np.random.default_rng(seed)
    va = _rig_view(_rodrigues(np.array([0.0, 0.02, 0.0])), np.array([0.0, 0.0, 0.0]), W, H, n_levels)
    vb = _rig_view(_rodrigues(np.array([0.01, -0.04, 0.005])), np.array([-0.6, 0.05, 0.1]), W, H, n_levels)
    sides, kps_all, nodes_all, has_all = [], [], [], []
    K = RIG_K.astype(np.float32)
    # points in front of camera `cam` of key frame 1
    R1 = va["Rsw"][cam].reshape(3, 3).astype(np.float64); t1 = va["tsw"][cam].astype(np.float64)
    R2 = vb["Rsw"][cam].reshape(3, 3).astype(np.float64); t2 = vb["tsw"][cam].astype(np.float64)
    uv = rng.uniform([10, 10], [W - 10, H - 10], (n, 2)); d = rng.uniform(3, 25, n)
    pc1 = np.stack([(uv[:, 0] - K[cam, 2]) / K[cam, 0] * d, (uv[:, 1] - K[cam, 3]) / K[cam, 1] * d, d], 1)
    Xw = (pc1 - t1) @ R1
    pc2 = Xw @ R2.T + t2
    uv2 = np.stack([K[cam, 0] * pc2[:, 0] / pc2[:, 2] + K[cam, 2], K[cam, 1] * pc2[:, 1] / pc2[:, 2] + K[cam, 3]], 1)
    node = rng.integers(0, n_nodes, n) * 5 + 2
    base = random_descriptors(seed * 41, n)
    out = {}
    for which, (uvk, nk) in enumerate(((uv, n), (uv2, n))):
        per_cam_k, per_cam_d, per_cam_nodes = [], [], []
        for c in range(2):
            if c == cam:
                m = nk + nk // 4
                k = np.zeros(m, KP_DTYPE)
                k["x"][:nk] = uvk[:, 0] + rng.normal(0, 0.7, nk); k["y"][:nk] = uvk[:, 1] + rng.normal(0, 0.7, nk)
                off = rng.random(nk) < 0.15                       # off the epipolar line
                k["y"][:nk][off] += rng.uniform(8, 40, off.sum())
                k["x"][nk:] = rng.uniform(0, W, m - nk); k["y"][nk:] = rng.uniform(0, H, m - nk)
                k["octave"] = rng.integers(0, n_levels, m)
                k["angle"][:nk] = ((40.0 if which == 0 else 10.0) + rng.normal(0, 3, nk)) % 360
                wild = rng.random(nk) < 0.1
                k["angle"][:nk][wild] = rng.uniform(0, 360, wild.sum())
                k["angle"][nk:] = rng.uniform(0, 360, m - nk)
                dd = np.concatenate([random_descriptors(seed * 43 + which, nk, p_flip, base), random_descriptors(seed * 47 + which, m - nk)])
                nn = np.concatenate([node, rng.integers(0, n_nodes + 4, m - nk) * 5 + 2])
                stray = rng.random(m) < 0.05
                nn[stray] = rng.integers(0, n_nodes, stray.sum()) * 5 + 2
                order = rng.permutation(m)                          # features are not stored in correspondence order
                k, dd, nn = k[order], dd[order], nn[order]
            else:
                m = 300
                k = np.zeros(m, KP_DTYPE)
                k["x"] = rng.uniform(0, W, m); k["y"] = rng.uniform(0, H, m); k["octave"] = rng.integers(0, n_levels, m); k["angle"] = rng.uniform(0, 360, m)
                dd = random_descriptors(seed * 53 + which + c, m); nn = rng.integers(0, n_nodes, m) * 5 + 2
            k["class_id"] = -1
            per_cam_k.append(k); per_cam_d.append(dd); per_cam_nodes.append(nn)
        kk = np.concatenate(per_cam_k)
        side = dict(n_kp=np.array([len(x) for x in per_cam_k], np.int32), desc=np.concatenate(per_cam_d), angle=kk["angle"].astype(np.float32), **_csr(per_cam_nodes))
        out[f"K{which + 1}"] = side; out[f"kps{which + 1}"] = kk
        out[f"has_mp{which + 1}"] = (rng.random(len(kk)) < 0.3).astype(np.uint8)
    # F12 = K1^-T [t12]x R12 K2^-1 with R12 = R1 R2^T, t12 = -R12 t2 + t1   (src/LocalMapping.cc ComputeF12)
    R12 = R1 @ R2.T; t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    Km = np.array([[K[cam, 0], 0, K[cam, 2]], [0, K[cam, 1], K[cam, 3]], [0, 0, 1]], np.float64)
    out["F12"] = (np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)).astype(np.float32).reshape(9)
    out["C1sw"] = va["Ow"][cam]; out["R2sw"] = vb["Rsw"][cam]; out["t2sw"] = vb["tsw"][cam]; out["K2cam"] = K[cam]
    out["scale_factors"] = scale_factors(n_levels); out["cam"] = cam
    return out


# ------------------------------------------------------------------------------------------------ vocabulary
def vocabulary(seed, k=10, L=3, p_flip=0.12, frac_stopped=0.02, ragged=False):
    """A synthetic DBoW2 vocabulary tree in the layout of ORBvoc.txt (rows in creation order of the hierarchical k-means: the k children
    of a node are created together, then each child is expanded): child descriptors are noisy copies of their parent's, leaf weights
    are idf-like positive doubles, a few words are stopped (weight 0).  ragged: some inner nodes have fewer than k children."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]

    def expand(pid, pdesc, level):
        nk = k if not ragged else int(rng.integers(2, k + 1))
        base = rng.integers(0, 256, (nk, 32), dtype=np.uint8) if level == 1 else random_descriptors(int(rng.integers(1 << 30)), nk, p_flip, np.tile(pdesc, (nk, 1)))
        ids = []
        for j in range(nk):
            ids.append(len(parent))
            parent.append(pid); desc.append(base[j])
            is_leaf = level == L
            leaf.append(1 if is_leaf else 0)
            weight.append(0.0 if (not is_leaf or rng.random() < frac_stopped) else float(rng.uniform(0.5, 9.0)))
        if level < L:
            for j, nid in enumerate(ids):
                expand(nid, base[j], level + 1)

    expand(0, None, 1)
    return dict(k=k, L=L, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8), desc=np.stack(desc), weight=np.array(weight, np.float64))


def vocabulary_text(voc):
    """the same vocabulary as the lines of an ORBvoc.txt-style file (TemplatedVocabulary::saveToTextFile)"""
    lines = [f"{voc['k']} {voc['L']} 0 0"]
    for i in range(1, len(voc["parent"])):
        lines.append(f"{voc['parent'][i]} {int(voc['is_leaf'][i])} " + " ".join(str(int(b)) for b in voc["desc"][i]) + f" {float(voc['weight'][i])!r}")
    return lines


def vocabulary_features(seed, voc, n, p_flip=0.06):
    """descriptors that are noisy copies of random leaf descriptors (so that words repeat inside an image)"""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(voc["is_leaf"])
    pick = leaves[rng.integers(0, max(len(leaves) // 3, 1), n)]
    return random_descriptors(seed + 1, n, p_flip, voc["desc"][pick])
