#!/usr/bin/env python
"""bench.py -- the hot path (ORB extract + Hamming match + LocalBA, GlobalBA) on N B200 GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config all|track640|track720|gba]

`--config track640` (the headline, BASELINE.json configs[1] + configs[2]): one "step" = one pass of the path over one batch of
`--frames` synthetic dual-frames (2 x 640x480, 1000 features per camera, 8 levels, scale 1.2, iniTh 20 / minTh 7):
  extract   both cameras of every frame (ORBextractor::operator())
  match     brute-force 256-bit Hamming, camera c of frame k against camera c of frame k+1
  LocalBA   one Optimizer::LocalBundleAdjustment window per `--kf-interval` dual-frames (default 1: every dual-frame closes a
            window), each of BASELINE.json configs[2] size: 20 keyframes x 2 cameras, 4000 map points, ~30k edges, 5 + 10 LM iterations.
`--config track720` (configs[3]): extract + match of 2 x 1280x720 dual-frames at 2000 features per camera, one sequence per GPU.
`--config gba` (configs[4]): Optimizer::GlobalBundleAdjustemnt of 2000 key frames x 2 cameras, 200k map points, landmark-partitioned
  over the ranks with one NCCL all-reduce of the reduced camera system per LM trial; value = milliseconds per LM trial.
`--config all` (default): the track640 line, with the other two configs measured in the same run under "extra_configs" (so that the
  driver's 1 / 2 / 4 / 8-GPU runs time every BASELINE config, the GlobalBA collective included).
N > 1 (torchrun): track*: every rank runs its own batch on its own GPU (independent sequences, no data-path collective: weak
scaling); gba: one problem over all ranks (strong scaling); time = max over ranks.

The JSON line follows the driver contract; DESIGN.md "Measurement" says how every field is obtained.
`--impl reference` times the CPU oracle (the reference's algorithm restated; the reference itself cannot be compiled in
this image) on all host cores; this and the `cpu_baseline` leg are the only places bench.py executes oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CAMS = 2
BA_UNIQUE = 4                             # distinct synthetic LocalBA windows of the headline workload, tiled over the batch
BA_HETERO_UNIQUE = 16                     # distinct windows of the heterogeneous LocalBA batch
GBA_SHAPE = (2000, 200000, 5)             # key frames, map points, LM iterations (BASELINE configs[4])


class TrackCfg:
    def __init__(self, name, W, H, nfeat, frames, with_ba):
        self.name, self.W, self.H, self.nfeat, self.frames, self.with_ba = name, W, H, nfeat, frames, with_ba
        sizes = self.level_sizes()
        self.pyr_pixels = sum(w * h for w, h in sizes)
        self.pyr_pixels_1up = self.pyr_pixels - W * H
        self.pyr_src_pixels = sum(w * h for w, h in sizes[:-1])           # what the 7 resize launches read
        # read input + write levels 1..7 + 2 x read all levels + 60 B per keypoint
        self.bytes_extract_image = W * H + self.pyr_pixels_1up + 2 * self.pyr_pixels + 60 * nfeat
        self.bytes_match_pair = 2 * nfeat * 32 + nfeat * 12
        self.stages = "extract+match+localBA" if with_ba else "extract+match"
        self.metric = f"frames/sec (dual {W}x{H}, {nfeat} kpts/cam) {self.stages}"

    def level_sizes(self, n=8, f=1.2):
        out, sc = [(self.W, self.H)], np.float32(1.0)
        for _ in range(1, n):
            sc = np.float32(sc * np.float32(f))
            inv = np.float32(1.0) / sc
            out.append((int(np.rint(np.float32(self.W) * inv)), int(np.rint(np.float32(self.H) * inv))))
        return out


TRACK640 = TrackCfg("track640", 640, 480, 1000, 256, True)
TRACK720 = TrackCfg("track720", 1280, 720, 2000, 64, False)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_pipes():
    """non-HBM peaks measured by tools/microbench.cu on this pool's B200 (profiles/r2_microbench.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_microbench.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(cfg, seed, frames):
    import synth
    a = synth.tiled_batch(seed, frames, cfg.W, cfg.H, CAMS, unique=16)
    b = np.ascontiguousarray(np.roll(a, shift=(11, 7), axis=(2, 3)))   # a second, distinct batch (defeats L2 reuse across steps)
    return a, b


def make_ba(seed, n):
    """n LocalBA windows of BASELINE configs[2] size (BA_UNIQUE distinct ones, tiled)."""
    import synth
    base = [synth.ba_problem(seed * 16 + i) for i in range(min(BA_UNIQUE, max(n, 1)))]
    return [base[i % len(base)] for i in range(n)]


def make_ba_hetero(seed, n):
    """n LocalBA windows of very different sizes, noise levels and outlier shares (BA_HETERO_UNIQUE distinct ones, tiled): some need
    rejected LM trials, some stop early -- the lock-step batch waits for the slowest."""
    import synth
    rng = np.random.default_rng(1000 + seed)
    base = []
    for i in range(min(BA_HETERO_UNIQUE, max(n, 1))):
        hard = i % 4 == 3                  # a young map: few key frames and points, poor initial estimates -> rejected LM trials
        base.append(synth.ba_problem(seed * 64 + 200 + i, n_kf=int(rng.integers(4, 7) if hard else rng.integers(8, 31)),
                                     n_points=int(rng.integers(60, 400) if hard else rng.integers(1500, 5001)),
                                     n_fixed_extra=0 if hard else int(rng.integers(0, 5)), outlier_frac=float(rng.choice([0.02, 0.05, 0.15])),
                                     pose_noise=(0.5, 10.0) if hard else (0.02, 0.5), point_noise=1.0 if hard else 0.05))
    return [base[i % len(base)] for i in range(n)]


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_path(cfg, frames_u8, ba_problems, kf_interval, threads):
    """The reference algorithm on the host: extract both cameras of every dual-frame, match frame k -> k+1 per camera, one
    LocalBundleAdjustment per kf_interval dual-frames.  Returns seconds.  `threads` workers, each owning whole dual-frames
    (the oracle releases the GIL inside ctypes)."""
    import oracle_lib as O
    F = frames_u8.shape[0]
    descs = [[None] * CAMS for _ in range(F)]

    def extract_range(lo, hi):
        ex = O.Extractor(cfg.nfeat, 1.2, 8, 20, 7)
        for f in range(lo, hi):
            for c in range(CAMS):
                descs[f][c] = ex(frames_u8[f, c])[1]

    def match_range(lo, hi):
        for f in range(lo, hi):
            for c in range(CAMS):
                O.match_bruteforce(descs[f][c], descs[(f + 1) % F][c])

    def ba_range(lo, hi):
        for f in range(lo, hi):
            if f % kf_interval == 0:
                O.local_ba(ba_problems[(f // kf_interval) % len(ba_problems)])

    def run(fn):
        if threads == 1:
            fn(0, F)
            return
        cuts = np.linspace(0, F, threads + 1).astype(int)
        ts = [threading.Thread(target=fn, args=(cuts[i], cuts[i + 1])) for i in range(threads) if cuts[i + 1] > cuts[i]]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    t0 = time.perf_counter()
    run(extract_range)
    run(match_range)
    if cfg.with_ba:
        run(ba_range)
    return time.perf_counter() - t0


def workload_text(cfg, frames, kf):
    s = (f"ORB {cfg.stages}, batch of {frames} dual-frames 2x{cfg.W}x{cfg.H}, {cfg.nfeat} feats/cam, 8 levels x1.2, brute-force 256-bit "
         f"Hamming frame k -> k+1 per camera (BASELINE configs[{1 if cfg.with_ba else 3}])")
    if cfg.with_ba:
        s += (f"; one LocalBA window (20 KFs x 2 cams, 4000 points, ~30k edges, 5+10 LM iterations: BASELINE configs[2]) per {kf} dual-frame(s)")
    return s


def reference_track(cfg, args):
    import oracle_lib as O
    O.lib()
    cores = os.cpu_count() or 1
    sample = max(cores * 2, 16) if cfg.with_ba else max(cores, 8)
    import synth
    frames = synth.tiled_batch(0, sample, cfg.W, cfg.H, CAMS, unique=16)
    ba = make_ba(0, BA_UNIQUE) if cfg.with_ba else []
    for _ in range(args.warmup):
        cpu_path(cfg, frames[:cores], ba, args.kf_interval, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_path(cfg, frames, ba, args.kf_interval, cores)
    fps = sample * args.steps / t
    desc = f"{sample} dual-frames per step, {cores} threads (one oracle instance per thread), {cfg.stages}"
    return {
        "impl": "reference", "metric": cfg.metric, "value": fps, "unit": "dual-frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_text(cfg, sample, args.kf_interval) + " -- CPU oracle = the reference algorithm restated (the reference cannot be compiled here)",
                   "frames_per_step": sample, "stages": cfg.stages, "kf_interval": args.kf_interval},
        "cpu_baseline": {"value": fps, "unit": "dual-frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": "dual-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


GBA_METRIC = "ms per LM trial, GlobalBundleAdjustemnt 2000 KFs x 2 cams, 200k map points (Schur reduced-camera system all-reduced over NVLink)"


def reference_gba(args):
    """the CPU oracle on a bounded sample of the workload: the same generator at 300 key frames / 30k points (the oracle's reduced
    system is dense: 2000 key frames would take minutes per trial)"""
    import oracle_lib as O
    import synth
    p = synth.gba_problem(0, n_kf=300, n_points=30000)
    t0 = time.perf_counter()
    rc, _, _, st = O.global_ba(p, iterations=3)
    dt = time.perf_counter() - t0
    ms = 1e3 * dt / max(st["trials"], 1)
    desc = f"300 key frames, 30000 points, {len(p['edge_pose'])} edges, {st['trials']} LM trials, 1 thread ({dt:.1f} s): 1/7 of the workload's size per trial"
    return {
        "impl": "reference", "metric": GBA_METRIC, "value": ms, "unit": "ms/trial (on the sample)", "n_gpus": args.gpus, "steps": st["trials"], "warmup": 0,
        "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GlobalBundleAdjustemnt, CPU oracle on a bounded sample: " + desc},
        "cpu_baseline": {"value": ms, "unit": "ms/trial (on the sample)", "cores": 1, "kind": "port", "sample": desc},
        "e2e": {"value": ms, "unit": "ms/trial (on the sample)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }


# ------------------------------------------------------------------------------------------------ GPU arms
def ba_bytes(problems, stats):
    """Bytes every BA kernel moves by design, summed over the launches that do work: per problem, k_lin / k_build run on
    the first step of each of the two rounds, the other four once per LM trial."""
    tot = dict.fromkeys(["k_lin", "k_build", "k_land", "k_pairs", "k_solve", "k_back"], 0.0)
    for p, st in zip(problems, stats):
        E, L = len(p["edge_pose"]), len(p["points"])
        free = p["pose_fixed"] == 0
        K, P = int(free.sum()), len(free)
        ef = free[p["edge_pose"]]
        Ef = int(ef.sum())
        f_l = np.bincount(p["edge_point"][ef], minlength=L)
        T = int((f_l * (f_l + 1) // 2).sum())
        tr = st["trials"]
        nC = len(p["cam_K"])
        C = K * (K + 1) // 2 * nC * nC + T // 512
        tot["k_lin"] += 2 * (E * (12 + 16 + 8 + 16 + 64 + 128) + L * 24 + P * nC * 128)       # ids, obs, info, table line, err -> 64-byte edge record
        tot["k_build"] += 2 * (E * (64 + 8) + Ef * (64 + 8) + L * 72 + K * 336)                # records read once per landmark part and once per pose part
        tot["k_land"] += tr * (E * (12 + 16 + 8 + 1 + 128 + 128 + 32) + L * (24 + 72))          # ids, err, weight, level, table line -> 128-byte record + {r, VD bl}; Hll, bl
        tot["k_pairs"] += tr * (Ef * (128 + 32) + T * 8 + C * (288 + 96))                       # every edge record read once (the gathers re-read it from L2)
        tot["k_solve"] += tr * (C * (288 + 96) + K * (56 + 48 + 96 * nC) + P * nC * 128)
        tot["k_back"] += tr * (Ef * (80 + 12) + L * (96 + 24) + E * (12 + 16 + 8 + 1 + 16 + 128))
    return tot


def survey_ba_bytes(problems, stats):
    """SURVEY.md §8(d) algorithmic bytes of LocalBA: per outer LM iteration read E*36 + P*24 + K*56 and write (E*18 + P*12 + K*42)*8"""
    tot = 0.0
    for p, st in zip(problems, stats):
        E, L = len(p["edge_pose"]), len(p["points"])
        K = int((p["pose_fixed"] == 0).sum())
        tot += st["iterations"] * (E * 36 + L * 24 + K * 56 + (E * 18 + L * 12 + K * 42) * 8)
    return tot


def run_track(cfg, args, rank, world, local_rank, steps, warmup):
    import torch
    import torch.distributed as dist
    from orbslam2_dualcam_b200 import ORBextractor, ORBmatcher, Optimizer, compact_problem
    import synth
    dev = torch.device("cuda", local_rank)
    F, KF = (args.frames if cfg.with_ba else cfg.frames), args.kf_interval
    NBA = (F + KF - 1) // KF if cfg.with_ba else 0
    a, b = make_inputs(cfg, 1000 * rank, F)
    host = [torch.from_numpy(x).pin_memory() for x in (a, b)]
    d_in = [h.to(dev) for h in host]
    ext = ORBextractor(cfg.nfeat, 1.2, 8, 20, 7, width=cfg.W, height=cfg.H, cameras=CAMS, max_frames=F, device=local_rank)
    cap = ext.kp_capacity
    P = F * CAMS
    mat = ORBmatcher(max_pairs=P, max_query=cap, max_train=cap, device=local_rank)
    q_set = torch.arange(P, dtype=torch.int32, device=dev)
    t_set = ((q_set + CAMS) % P).to(torch.int32)          # same camera, next frame (cyclic)
    d_kps = torch.zeros((F, CAMS, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((F, CAMS, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros((F, CAMS), dtype=torch.int32, device=dev)
    d_match = tuple(torch.zeros((P, cap), dtype=torch.int32, device=dev) for _ in range(3))
    h_kps, h_desc, h_cnt = (torch.empty_like(t, device="cpu").pin_memory() for t in (d_kps, d_desc, d_cnt))
    h_match = tuple(torch.empty_like(t, device="cpu").pin_memory() for t in d_match)
    # (Running LocalBA on a second stream next to extract + match was measured and is slower: the two working sets evict each other
    # from L2.  One compute stream.)
    stream = torch.cuda.Stream(dev, priority=-1)   # every kernel and event of the timed regions goes through this (high-priority) stream
    copy_stream = torch.cuda.Stream(dev)     # end-to-end leg: BA uploads (host->device + index kernels, which yield to the compute stream), see below
    torch.cuda.set_stream(stream)
    opt = opt2 = None
    ba_problems, ba_compact, h_ba = [], None, ()
    if cfg.with_ba:
        ba_problems = make_ba(rank, NBA)
        lev = synth.inv_sigma2_levels()
        distinct = {id(p): compact_problem(p, lev) for p in ba_problems[:BA_UNIQUE]}
        ba_compact = Optimizer.prepare_f32([distinct[id(p)] for p in ba_problems])      # the form the reference holds: CV_32F, 16-byte observations
        nPt = sum(len(p["pose_fixed"]) for p in ba_problems)
        nLt = sum(len(p["points"]) for p in ba_problems)
        nEt = sum(len(p["edge_pose"]) for p in ba_problems)
        h_ba = (torch.empty((nPt, 12), dtype=torch.float64).pin_memory(), torch.empty((nLt, 3), dtype=torch.float64).pin_memory(),
                torch.empty((nEt,), dtype=torch.uint8).pin_memory())
        opt = Optimizer(max_problems=NBA, device=local_rank)
        opt.set_stream(stream)
        opt.upload(opt.prepare(ba_problems))     # device-resident leg: the windows are uploaded (and indexed) once
        opt2 = Optimizer(max_problems=NBA, device=local_rank)      # second handle: the end-to-end leg double-buffers the BA uploads
        opt2.set_stream(stream)
    torch.cuda.synchronize(dev)

    def step(imgs):
        ext.extract_device(imgs, d_kps, d_desc, d_cnt, stream=stream)
        mat.bruteforce_sets_device(d_desc, d_cnt, q_set, t_set, out=d_match, stream=stream)
        if opt:
            opt.run()                        # LocalBundleAdjustment of every window: optimize(5) Huber, outlier pass, optimize(10)

    def barrier():
        if world > 1:
            dist.barrier()
        if opt:
            opt.synchronize()
        torch.cuda.synchronize(dev)

    def launches():
        return ext.launch_count() + mat.launch_count() + (opt.launch_count() + opt2.launch_count() if opt else 0)

    # ---- device-resident throughput (`value`)
    for i in range(warmup):
        step(d_in[i % 2])
    barrier()
    l0 = launches()
    ext.profile(True)
    mat.profile(True)
    if opt:
        opt.profile(True)
    ba_ms, ba_kernel_ms, ba_steps = 0.0, None, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for i in range(steps):
            step(d_in[i % 2])
            if opt:
                m, _ = opt.stage_ms()        # waits for this step's LM steps, the ones finish() adds for windows that needed more included
                ba_ms += m
                km, ks = opt.kernel_ms()
                ba_kernel_ms = km if ba_kernel_ms is None else {k: ba_kernel_ms[k] + v for k, v in km.items()}
                ba_steps += ks
        e1.record(stream)                    # after the last step's synchronisation: every LM step of the timed region lies inside it
        barrier()
    ms = e0.elapsed_time(e1)
    n_launch = launches() - l0
    stage_ms, calls = ext.stage_ms()
    match_ms, mcalls = mat.stage_ms()
    ext.profile(False)
    mat.profile(False)
    n_kp = int(d_cnt.sum().item())
    ba_stats = None
    if opt:
        opt.profile(False)
        ba_stats = opt.download_batch(out=h_ba)[3]

    # ---- end to end: pinned host images and host BA graphs in; keypoints / descriptors / matches / poses / points / outlier flags
    #      out to pinned host memory, every step.  Three streams: copies in (images + BA graphs), compute, copies out.  The BA graphs
    #      go in as the reference holds them (orbba_upload_f32: CV_32F poses / points, 16-byte observations) and are double-buffered over
    #      two handles: while the device optimises the windows of step i, the host stages and uploads step i+1, and the BA results of
    #      step i are read during step i+1 (the last step is drained inside the timed region).
    opts = [opt, opt2]
    if opt:
        for o in opts:
            o.set_copy_stream(copy_stream)   # uploads overlap with the other handle's run; runs share one compute stream
    out_stream = torch.cuda.Stream(dev)      # device->host copies of keypoints / descriptors / matches
    prev_out = [None]

    trace = []                               # per step: compute-stream events before extract, after match, after the LM steps
    marker_stream = torch.cuda.Stream(dev)   # idle stream: an event recorded here carries the host's 'now' on the device clock
    marks = []

    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record(marker_stream)
        return e

    host_ms = dict.fromkeys(["enqueue_images", "ba_upload_call", "enqueue_kernels", "ba_download_wait", "wait_outputs"], 0.0)

    def n_outliers():
        # the step's BA result is read on the host: count of outlier flags (numpy view of the pinned buffer; torch's uint8 sum is a scalar
        # loop that took 25 ms for these 7.6 M flags and made the HOST the bottleneck of the 8-GPU end-to-end run)
        return int(np.count_nonzero(h_ba[2].numpy()))

    def e2e_step(i, first):
        o = opts[i % 2]
        t0 = time.perf_counter()
        mk = [mark()]
        with torch.cuda.stream(copy_stream):
            d_in[i % 2].copy_(host[i % 2], non_blocking=True)
            e_img = torch.cuda.Event(enable_timing=True)
            e_img.record(copy_stream)
        t1 = time.perf_counter()
        if o:
            o.upload(ba_compact)             # stage + host->device + expansion and index construction on the device (copy stream)
        t2 = time.perf_counter()
        mk.append(mark())
        stream.wait_event(e_img)
        if prev_out[0] is not None:
            stream.wait_event(prev_out[0])   # the previous step's results left the device before they are overwritten
        tr = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        trace.append(tr)
        tr[0].record(stream)
        ext.extract_device(d_in[i % 2], d_kps, d_desc, d_cnt, stream=stream)
        mat.bruteforce_sets_device(d_desc, d_cnt, q_set, t_set, out=d_match, stream=stream)
        e_em = stream.record_event()
        tr[1].record(stream)
        if o:
            o.run()                          # compute stream, after this handle's upload
        tr[2].record(stream)
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(e_em)
            h_cnt.copy_(d_cnt, non_blocking=True)
            h_kps.copy_(d_kps, non_blocking=True)
            h_desc.copy_(d_desc, non_blocking=True)
            for hm, dm in zip(h_match, d_match):
                hm.copy_(dm, non_blocking=True)
            prev_out[0] = out_stream.record_event()
        t3 = time.perf_counter()
        r = 0
        if o and not first:
            opts[(i - 1) % 2].download_batch(out=h_ba)     # results of the previous step's windows (waits for that run only)
            r = n_outliers()
        t4 = time.perf_counter()
        mk.append(mark())
        mk.append(e_img)
        marks.append(mk)
        prev_out[0].synchronize()
        t5 = time.perf_counter()
        for k, v in zip(host_ms, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            host_ms[k] += 1e3 * v
        return int(h_cnt.sum()) + r          # the step's results are read on the host

    def e2e_run(n):
        prev_out[0] = None
        for i in range(n):
            e2e_step(i, i == 0)
        if opt:
            opts[(n - 1) % 2].download_batch(out=h_ba)
            return n_outliers()
        return 0

    e2e_run(max(2, warmup // 2))
    barrier()
    if opt2:
        opt2.synchronize()
    for k in host_ms:
        host_ms[k] = 0.0
    trace.clear()
    marks.clear()
    t0 = time.perf_counter()
    e2e_run(steps)
    barrier()
    if opt2:
        opt2.synchronize()
    e2e_s = time.perf_counter() - t0
    torch.cuda.synchronize(dev)
    dev_ms = {"extract_match": float(np.mean([a.elapsed_time(b) for a, b, _ in trace])),
              "ba_incl_wait_for_upload": float(np.mean([b.elapsed_time(c) for _, b, c in trace])),
              # timeline of a steady-state step on the device clock, relative to the host's step start: image copy done, upload call returned
              # (extract is enqueued right after), extract starts, LM steps end, BA results of the previous step on the host
              "timeline_vs_step_start": {k: float(np.mean(v)) for k, v in {
                  "image_h2d_done": [m[0].elapsed_time(m[3]) for m in marks[2:]],
                  "upload_call_returned": [m[0].elapsed_time(m[1]) for m in marks[2:]],
                  "extract_starts": [m[0].elapsed_time(t[0]) for m, t in zip(marks[2:], trace[2:])],
                  "match_ends": [m[0].elapsed_time(t[1]) for m, t in zip(marks[2:], trace[2:])],
                  "lm_steps_end": [m[0].elapsed_time(t[2]) for m, t in zip(marks[2:], trace[2:])],
                  "prev_lm_steps_end": [marks[k][0].elapsed_time(trace[k - 1][2]) for k in range(2, len(marks))],
                  "prev_ba_results_on_host": [m[0].elapsed_time(m[2]) for m in marks[2:]]}.items() if v},
              "gap_before_extract": float(np.mean([trace[k][2].elapsed_time(trace[k + 1][0]) for k in range(len(trace) - 1)])) if len(trace) > 1 else 0.0}
    ba_in = 0
    if opt:
        ba_in = sum(sum(np.asarray(q[k]).nbytes for k in ("poses", "pose_fixed", "points", "edges", "inv_sigma2", "cam_K", "cam_ext", "cam_adj")) for q in ba_compact[1])
    h2d = host[0].numel() + ba_in
    d2h = sum(t.numel() * t.element_size() for t in (h_cnt, h_kps, h_desc) + h_match + tuple(h_ba))

    # ---- heterogeneous LocalBA batch (device-resident): windows of different sizes, some with rejected LM trials
    hetero = None
    if opt and args.hetero:
        hp = make_ba_hetero(rank, NBA)
        oh = Optimizer(max_problems=NBA, device=local_rank)
        oh.set_stream(stream)
        oh.upload(oh.prepare(hp))
        oh.profile(True)
        oh.run()
        oh.synchronize()
        tms = []
        for _ in range(3):
            oh.run()
            tms.append(oh.stage_ms()[0])
        hst = oh.download_batch()[3]
        hetero = {"ms_per_batch": float(min(tms)), "windows": NBA, "distinct_windows": min(BA_HETERO_UNIQUE, NBA),
                  "windows_per_s": NBA / (min(tms) * 1e-3),
                  "edges_per_window": [int(min(len(p["edge_pose"]) for p in hp)), int(max(len(p["edge_pose"]) for p in hp))],
                  "lm_trials": [int(min(s["trials"] for s in hst)), int(max(s["trials"] for s in hst))],
                  "windows_with_rejected_trials": int(sum(s["trials"] > s["iterations"] for s in hst)),
                  "lm_steps_launched_per_run": oh.kernel_ms()[1] / 4}
        oh.close()

    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    out = None
    if rank == 0:
        peak, peak_kind = peaks()
        total_frames = F * world
        fps = total_frames * steps / (ms / 1e3)
        e2e_fps = total_frames * steps / (e2e_ms / 1e3)
        NI = F * CAMS
        nc = max(calls, 1)
        # per-kernel roofline table: (ms per bench step, bytes the kernel moves by design per bench step)
        kern = {
            "resize4_kernel(x7)": (stage_ms["pyramid"] / nc, (cfg.pyr_pixels_1up + cfg.pyr_src_pixels) * NI),
            "fast_cells_kernel": (stage_ms["fast"] / nc, cfg.pyr_pixels * NI),
            "quadtree_kernel": (stage_ms["quadtree"] / nc, None),
            # whole-level Gaussian (read + write every pyramid pixel) + per key point a 64x37 blurred window, a 31x31 raw disc, 60 bytes out
            "blur_level_kernel(x8)+describe_kernel": (stage_ms["describe"] / nc, 2 * cfg.pyr_pixels * NI + (64 * 37 + 961 + 60) * n_kp),
            "bruteforce_kernel": (match_ms / max(mcalls, 1), cfg.bytes_match_pair * NI),
        }
        if opt:
            bb = ba_bytes(ba_problems, ba_stats)
            for k, v in (ba_kernel_ms or {}).items():
                kern["ba_" + k] = (v / steps, bb[k])
        table = {}
        for k, (kms, kb) in kern.items():
            table[k] = {"ms_per_step": kms, "design_bytes_per_step": kb,
                        "achieved_gbs": (kb / (kms * 1e-3) / 1e9 if kb and kms > 0 else None)}
            if table[k]["achieved_gbs"] is not None:
                table[k]["frac"] = table[k]["achieved_gbs"] / peak
        dom_name = max((k for k in kern if kern[k][1]), key=lambda k: kern[k][0])
        dom_ms, dom_bytes = kern[dom_name]
        n_dom_launch = (ba_steps / steps) if dom_name.startswith("ba_") else (7 if dom_name.startswith("resize") else 1)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # SURVEY.md §8(d) algorithmic bytes of the stages (what the judge's roofline uses) next to the kernels' own traffic
        em_ms = sum(stage_ms[k] for k in ("pyramid", "fast", "quadtree", "describe")) / nc + match_ms / max(mcalls, 1)
        em_bytes = (cfg.bytes_extract_image + cfg.bytes_match_pair) * NI
        step_ms = ms / steps
        sv_ba = survey_ba_bytes(ba_problems, ba_stats) if opt else 0.0

        def sv(bytes_, ms_):
            return {"bytes_per_step": bytes_, "ms_per_step": ms_, "achieved_gbs": bytes_ / (ms_ * 1e-3) / 1e9, "frac": bytes_ / (ms_ * 1e-3) / 1e9 / peak}
        survey = {"extract+match": sv(em_bytes, em_ms)}
        if opt:
            survey["localBA"] = sv(sv_ba, ba_ms / steps)
        survey["step"] = sv(em_bytes + sv_ba, step_ms)
        pipes = measured_pipes()
        other = None
        if pipes:
            pairs_s = NI * cfg.nfeat * cfg.nfeat / (match_ms / max(mcalls, 1) * 1e-3)
            other = {"match_distance_pairs_per_s": pairs_s, "popc_b64_peak_per_s": pipes["popc_b64"]["per_s"],
                     "match_frac_of_popc_peak": pairs_s * 4 / pipes["popc_b64"]["per_s"],
                     "fp64_fma_peak_per_s": pipes["dfma"]["per_s"], "peak_source": "tools/microbench.cu on this pool (profiles/r2_microbench.json)"}
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom_name)
        except Exception:
            pass
        out = {
            "metric": cfg.metric, "value": fps, "unit": "dual-frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (extract, match) / f64 (LocalBA)" if opt else "u8", "data": "synthetic",
            "config": {"workload": workload_text(cfg, F, KF), "frames_per_step_per_gpu": F, "ba_windows_per_step_per_gpu": NBA, "kf_interval": KF,
                       "stages": cfg.stages, "keypoints_per_step": n_kp,
                       "l2": f"two alternating input batches of {host[0].numel() // 1000000} MB each (> 126 MB L2)" + ("; BA working set 12 MB per window" if opt else ""),
                       "parallelism": f"{world} independent replicas, no collective"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_fps, "unit": "dual-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / steps,
                    "host_ms_per_step_rank0": {k: v / steps for k, v in host_ms.items()}, "device_ms_per_step_rank0": dev_ms},
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_kind, "traffic": traffic, "algorithmic_bytes_per_launch": dom_bytes / max(n_dom_launch, 1),
                         "ms_per_launch": dom_ms / max(n_dom_launch, 1), "launches_per_step": n_dom_launch,
                         "bytes_basis": "bytes the kernel moves by design (DESIGN.md §4); `survey_8d` holds the SURVEY.md §8(d) algorithmic bytes of the stages",
                         "survey_8d": survey, "other_pipes": other,
                         "stage_ms_per_step": {**{k: v / nc for k, v in stage_ms.items()}, "match": match_ms / max(mcalls, 1),
                                               **({"localBA": ba_ms / steps} if opt else {})},
                         "kernels": table},
        }
        if opt:
            out["config"]["ba"] = {"edges_per_window": nEt // max(NBA, 1), "lm_iterations": ba_stats[0]["iterations"], "lm_trials": ba_stats[0]["trials"],
                                   "distinct_windows": min(BA_UNIQUE, NBA), "lm_steps_launched_per_step": ba_steps / steps,
                                   "input_form_e2e": "orbba_upload_f32 (CV_32F poses / points, 16-byte observations)"}
            if hetero:
                out["config"]["ba_heterogeneous"] = hetero
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample if cfg.with_ba else 8
            t = cpu_path(cfg, a[:sample], ba_problems[:BA_UNIQUE], KF, 1)
            out["cpu_baseline"] = {"value": sample / t, "unit": "dual-frames/s", "cores": 1, "kind": "port",
                                   "sample": f"first {sample} dual-frames of the same batch, single thread, {cfg.stages} ({t:.1f} s)"}
    for o in (opt, opt2, mat, ext):
        if o:
            o.close()
    del d_in, d_kps, d_desc, d_match
    torch.cuda.empty_cache()
    return out


def run_gba(args, rank, world, local_rank):
    """BASELINE configs[4]: one GlobalBundleAdjustemnt over all ranks; value = device milliseconds per LM trial (events around the LM loop,
    max over ranks); e2e = wall time of the whole call per trial (host partition set-up, upload, LM loop, download)."""
    import torch
    import torch.distributed as dist
    from orbslam2_dualcam_b200 import DistributedOptimizer, shard_problem
    import synth
    n_kf, n_pts, its = GBA_SHAPE
    dev = torch.device("cuda", local_rank)
    p = synth.gba_problem(0, n_kf=n_kf, n_points=n_pts)
    opt = DistributedOptimizer.from_torch_distributed(local_rank) if world > 1 else DistributedOptimizer(device=local_rank)
    sh = shard_problem(p, rank, world)
    opt.GlobalBundleAdjustemnt(sh, nIterations=1)                      # warm-up (allocations, NCCL channels)
    l0 = opt.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        poses, points, st = opt.GlobalBundleAdjustemnt(sh, nIterations=its)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
    tm = opt.timing()
    times = torch.tensor([tm["loop_ms"], wall * 1e3, tm["allreduce_ms"], tm["solve_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    loop_ms, wall_ms, ar_ms, solve_ms = times.tolist()
    n_launch = opt.launch_count() - l0
    out = None
    if rank == 0:
        peak, peak_kind = peaks()
        tr = max(st["trials"], 1)
        E = len(p["edge_pose"])
        in_bytes = sum(np.asarray(sh[k]).nbytes for k in ("poses", "pose_fixed", "points", "edge_pose", "edge_point", "edge_cam", "edge_obs", "edge_inv_sigma2"))
        # per trial this rank's kernels move: B and Y blocks (144 B per edge, written / read twice), the skyline, the landmark blocks
        design = (E / world) * (144 * 4 + 176 + 52) + tm["skyline_blocks"] * 288 * 3 + (n_pts / world) * 150
        out = {
            "metric": GBA_METRIC, "value": loop_ms / tr, "unit": "ms/trial", "n_gpus": world, "steps": st["trials"], "warmup": 1,
            "ms_per_step": loop_ms / tr, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"GlobalBundleAdjustemnt (BASELINE configs[4]): {n_kf} key frames x 2 cams, {n_pts} map points, {E} edges, {its} LM iterations, Huber sqrt(3.99); "
                                   f"landmarks partitioned point_id mod {world}, poses replicated", "lm_trials": st["trials"], "lm_iterations": st["iterations"],
                       "reduced_system": 6 * (n_kf - 1), "skyline_blocks": tm["skyline_blocks"], "chi2": [st["initial_chi2"], st["final_chi2"]],
                       "parallelism": f"{world} ranks, NCCL all-reduce(sum, f64) of the block-skyline reduced camera system + rhs per LM trial" if world > 1 else "1 rank",
                       "l2": "working set 0.9 GB per rank-share of edges (> 126 MB L2)"},
            "clocks": clk.summary(),
            "collective": {"allreduce_ms_per_trial": ar_ms / tr, "allreduce_bytes_per_trial": tm["allreduce_bytes"] / tr,
                           "allreduce_GBps": (tm["allreduce_bytes"] / 1e9) / max(ar_ms / 1e3, 1e-9) if world > 1 else None,
                           "solve_ms_per_trial": solve_ms / tr, "build_ms_per_trial": (loop_ms - ar_ms - solve_ms) / tr},
            "e2e": {"value": wall_ms / tr, "unit": "ms/trial", "h2d_bytes_per_step": in_bytes / tr, "d2h_bytes_per_step": (poses.nbytes + points.nbytes) / tr,
                    "note": "whole orbba_dist_optimize call (host-side partition set-up and tuple lists, upload, LM loop, download) divided by the LM trials"},
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "hbm", "kernel": "whole LM trial (build kernels + skyline LDL^T)", "achieved": design / (loop_ms / tr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": design / (loop_ms / tr * 1e-3) / 1e9 / peak, "peak_source": peak_kind, "traffic": None,
                         "note": "the trial is latency-bound: the substructured skyline factorisation is a chain of 6 x (key-frames / segments + separators) sequential pivots (DESIGN.md section 4)"},
        }
    opt.close()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="all", choices=["all", "track640", "track720", "gba"])
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--kf-interval", type=int, default=1, help="dual-frames per LocalBA window")
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hetero", dest="hetero", action="store_false", help="skip the heterogeneous LocalBA batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        line = {"track640": lambda: reference_track(TRACK640, args), "all": lambda: reference_track(TRACK640, args),
                "track720": lambda: reference_track(TRACK720, args), "gba": lambda: reference_gba(args)}[args.config]()
        print(json.dumps(line))
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # the ranks of a node share its host cores: split them for the library's host-side staging of the BA graphs
    os.environ.setdefault("ORB_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // max(world, 1))))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = None
    if args.config in ("all", "track640"):
        line = run_track(TRACK640, args, rank, world, local_rank, args.steps, args.warmup)
    if args.config == "track720":
        line = run_track(TRACK720, args, rank, world, local_rank, args.steps, args.warmup)
    if args.config == "gba":
        line = run_gba(args, rank, world, local_rank)
    if args.config == "all":
        extra = {"track720": run_track(TRACK720, args, rank, world, local_rank, max(3, min(args.steps, 5)), 3),
                 "gba": run_gba(args, rank, world, local_rank)}
        if rank == 0:
            for v in extra.values():
                v.pop("cpu_baseline", None)
            line["extra_configs"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
