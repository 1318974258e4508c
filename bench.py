#!/usr/bin/env python
"""bench.py -- dual-frames/s of the hot path (ORB extract + brute-force match + LocalBA) on N B200 GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames 256] [--kf-interval 1]

One "step" = one pass of the path over one batch of `--frames` synthetic dual-frames (2 x 640x480, 1000 features per
camera, 8 levels, scale 1.2, iniTh 20 / minTh 7: BASELINE.json configs[1]):
  extract   both cameras of every frame (ORBextractor::operator())
  match     brute-force 256-bit Hamming, camera c of frame k against camera c of frame k+1
  LocalBA   one Optimizer::LocalBundleAdjustment window per `--kf-interval` dual-frames (default 1: every dual-frame closes a
            window), each of BASELINE.json configs[2] size: 20 keyframes x 2 cameras, 4000 map points, ~30k edges, 5 + 10 LM iterations.
N > 1 (torchrun): every rank runs its own batch on its own GPU (independent sequences, no data-path collective: weak
scaling); time = max over ranks.

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

W, H, CAMS, NFEAT = 640, 480, 2, 1000
METRIC = "frames/sec (dual 640x480, 1000 kpts/cam) extract+match+localBA"
STAGES = "extract+match+localBA"
# algorithmic bytes, SURVEY.md §8(d) / DESIGN.md §4
PYR_PIXELS = 950532                       # sum of the 8 level sizes of a 640x480 image
PYR_PIXELS_1UP = PYR_PIXELS - W * H       # levels 1..7
BYTES_EXTRACT_IMAGE = 2911596             # read input + write levels 1..7 + 2 x read all levels + 60 B per keypoint
BYTES_MATCH_PAIR = 76000                  # (nq + nt) x 32 + nq x 12
BA_UNIQUE = 4                             # distinct synthetic LocalBA windows, tiled over the batch


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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


def make_inputs(seed, frames):
    from orbslam2_dualcam_b200 import synth
    a = synth.tiled_batch(seed, frames, W, H, CAMS, unique=16)
    b = np.ascontiguousarray(np.roll(a, shift=(11, 7), axis=(2, 3)))   # a second, distinct batch (defeats L2 reuse across steps)
    return a, b


def make_ba(seed, n):
    """n LocalBA windows of BASELINE configs[2] size (BA_UNIQUE distinct ones, tiled)."""
    from orbslam2_dualcam_b200 import synth
    base = [synth.ba_problem(seed * 16 + i) for i in range(min(BA_UNIQUE, max(n, 1)))]
    return [base[i % len(base)] for i in range(n)]


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_path(frames_u8, ba_problems, kf_interval, threads):
    """The reference algorithm on the host: extract both cameras of every dual-frame, match frame k -> k+1 per camera, one
    LocalBundleAdjustment per kf_interval dual-frames.  Returns seconds.  `threads` workers, each owning whole dual-frames
    (the oracle releases the GIL inside ctypes)."""
    import oracle_lib as O
    F = frames_u8.shape[0]
    descs = [[None] * CAMS for _ in range(F)]

    def extract_range(lo, hi):
        ex = O.Extractor(NFEAT, 1.2, 8, 20, 7)
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
    run(ba_range)
    return time.perf_counter() - t0


def workload_text(frames, kf):
    return (f"ORB extract + match + LocalBA, batch of {frames} dual-frames 2x{W}x{H}, {NFEAT} feats/cam, 8 levels x1.2, brute-force 256-bit "
            f"Hamming frame k -> k+1 per camera (BASELINE configs[1]); one LocalBA window (20 KFs x 2 cams, 4000 points, ~30k edges, 5+10 LM "
            f"iterations: BASELINE configs[2]) per {kf} dual-frame(s)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle_lib as O
    O.lib()
    cores = os.cpu_count() or 1
    sample = max(cores * 2, 16)
    from orbslam2_dualcam_b200 import synth
    frames = synth.tiled_batch(0, sample, W, H, CAMS, unique=16)
    ba = make_ba(0, BA_UNIQUE)
    for _ in range(args.warmup):
        cpu_path(frames[:cores], ba, args.kf_interval, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_path(frames, ba, args.kf_interval, cores)
    fps = sample * args.steps / t
    desc = f"{sample} dual-frames per step, {cores} threads (one oracle instance per thread), {STAGES}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "dual-frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_text(sample, args.kf_interval) + " -- CPU oracle = the reference algorithm restated (the reference cannot be compiled here)",
                   "frames_per_step": sample, "stages": STAGES, "kf_interval": args.kf_interval},
        "cpu_baseline": {"value": fps, "unit": "dual-frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": "dual-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def ba_bytes(problems, stats):
    """Algorithmic bytes of every BA kernel summed over the launches that do work: per problem, k_lin / k_build run
    once per LM iteration, the other four once per LM trial."""
    tot = dict.fromkeys(["k_lin", "k_build", "k_land", "k_pairs", "k_solve", "k_back"], 0.0)
    for p, st in zip(problems, stats):
        E, L = len(p["edge_pose"]), len(p["points"])
        free = p["pose_fixed"] == 0
        K, P = int(free.sum()), len(free)
        ef = free[p["edge_pose"]]
        Ef = int(ef.sum())
        f_l = np.bincount(p["edge_point"][ef], minlength=L)
        T = int((f_l * (f_l + 1) // 2).sum())
        C = K * (K + 1) // 2 + T // 256
        it, tr = st["iterations"], st["trials"]
        nC = len(p["cam_K"])
        C = K * (K + 1) // 2 * nC * nC + T // 256
        tot["k_lin"] += it * (E * (12 + 16 + 8 + 16 + 64) + L * 24 + P * 56)                 # ids, obs, info, err -> 64-byte edge record
        tot["k_build"] += it * (E * (64 + 8) + Ef * (64 + 8) + L * 72 + K * 336)              # records read once per landmark part and once per pose part
        tot["k_land"] += tr * (Ef * (64 + 12 + 128 + 48) + L * 72)                        # k_trial: record -> {X Y Z 1/Z V VD} (128 B), v
        tot["k_pairs"] += tr * (Ef * (128 + 48) + T * 8 + C * 288 + K * nC * 48)              # every 128-byte edge record read once
        tot["k_solve"] += tr * (C * 288 + K * (336 + 104) + K * nC * 96)
        tot["k_back"] += tr * (Ef * (64 + 12) + L * (96 + 24) + E * (36 + 16))
    return tot


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from orbslam2_dualcam_b200 import ORBextractor, ORBmatcher, Optimizer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F, KF = args.frames, args.kf_interval
    NBA = (F + KF - 1) // KF
    a, b = make_inputs(1000 * rank, F)
    host = [torch.from_numpy(x).pin_memory() for x in (a, b)]
    d_in = [h.to(dev) for h in host]
    ba_problems = make_ba(rank, NBA)
    ext = ORBextractor(NFEAT, 1.2, 8, 20, 7, width=W, height=H, cameras=CAMS, max_frames=F, device=local_rank)
    cap = ext.kp_capacity
    P = F * CAMS
    mat = ORBmatcher(max_pairs=P, max_query=cap, max_train=cap, device=local_rank)
    opt = Optimizer(max_problems=NBA, device=local_rank)
    ba_prepared = opt.prepare(ba_problems)
    q_set = torch.arange(P, dtype=torch.int32, device=dev)
    t_set = ((q_set + CAMS) % P).to(torch.int32)          # same camera, next frame (cyclic)
    d_kps = torch.zeros((F, CAMS, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((F, CAMS, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros((F, CAMS), dtype=torch.int32, device=dev)
    d_match = tuple(torch.zeros((P, cap), dtype=torch.int32, device=dev) for _ in range(3))
    h_kps, h_desc, h_cnt = (torch.empty_like(t, device="cpu").pin_memory() for t in (d_kps, d_desc, d_cnt))
    h_match = tuple(torch.empty_like(t, device="cpu").pin_memory() for t in d_match)
    nPt = sum(len(p["pose_fixed"]) for p in ba_problems)
    nLt = sum(len(p["points"]) for p in ba_problems)
    nEt = sum(len(p["edge_pose"]) for p in ba_problems)
    h_ba = (torch.empty((nPt, 12), dtype=torch.float64).pin_memory(), torch.empty((nLt, 3), dtype=torch.float64).pin_memory(),
            torch.empty((nEt,), dtype=torch.uint8).pin_memory())
    # (Running LocalBA on a second stream next to extract + match was measured and is slower: 124 ms vs 81 ms per step, the two
    # working sets evict each other from L2.  One compute stream.)
    stream = torch.cuda.Stream(dev, priority=-1)   # every kernel and event of the timed regions goes through this (high-priority) stream
    copy_stream = torch.cuda.Stream(dev)     # end-to-end leg: BA uploads (host->device + index kernels, which yield to the compute stream), see below
    torch.cuda.set_stream(stream)
    opt.set_stream(stream)
    opt.upload(ba_prepared)                  # device-resident leg: the windows are uploaded (and indexed) once
    opt2 = Optimizer(max_problems=NBA, device=local_rank)      # second handle: the end-to-end leg double-buffers the BA uploads
    opt2.set_stream(stream)
    torch.cuda.synchronize(dev)

    def step(imgs):
        ext.extract_device(imgs, d_kps, d_desc, d_cnt, stream=stream)
        mat.bruteforce_sets_device(d_desc, d_cnt, q_set, t_set, out=d_match, stream=stream)
        opt.run()                            # LocalBundleAdjustment of every window: optimize(5) Huber, outlier pass, optimize(10)

    def barrier():
        if world > 1:
            dist.barrier()
        opt.synchronize()
        torch.cuda.synchronize(dev)

    def launches():
        return ext.launch_count() + mat.launch_count() + opt.launch_count() + opt2.launch_count()

    # ---- device-resident throughput (`value`)
    for i in range(args.warmup):
        step(d_in[i % 2])
    barrier()
    l0 = launches()
    ext.profile(True)
    mat.profile(True)
    opt.profile(True)
    ba_ms, ba_kernel_ms, ba_steps = 0.0, None, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step(d_in[i % 2])
            if i == args.steps - 1:
                e1.record(stream)
            m, _ = opt.stage_ms()            # waits for this step's LM steps (the device never idles more than the launch gap)
            ba_ms += m
            km, ks = opt.kernel_ms()
            ba_kernel_ms = km if ba_kernel_ms is None else {k: ba_kernel_ms[k] + v for k, v in km.items()}
            ba_steps += ks
        barrier()
    ms = e0.elapsed_time(e1)
    n_launch = launches() - l0
    stage_ms, calls = ext.stage_ms()
    match_ms, mcalls = mat.stage_ms()
    ext.profile(False)
    mat.profile(False)
    opt.profile(False)
    n_kp = int(d_cnt.sum().item())
    ba_stats = opt.download_batch(out=h_ba)[3]

    # ---- end to end: pinned host images and host BA graphs in; keypoints / descriptors / matches / poses / points / outlier flags
    #      out to pinned host memory, every step.  Three streams: copies in (images + BA graphs), compute, copies out.  The BA graphs
    #      are double-buffered over two handles: while the device optimises the windows of step i, the host flattens and uploads
    #      step i+1, and the BA results of step i are read during step i+1 (the last step is drained inside the timed region).
    opts = [opt, opt2]
    for o in opts:
        o.set_copy_stream(copy_stream)       # uploads overlap with the other handle's run; runs share one compute stream
    out_stream = torch.cuda.Stream(dev)      # device->host copies of keypoints / descriptors / matches
    prev_out = [None]

    def e2e_step(i, first):
        o = opts[i % 2]
        with torch.cuda.stream(copy_stream):
            d_in[i % 2].copy_(host[i % 2], non_blocking=True)
            e_img = copy_stream.record_event()
        o.upload(ba_prepared)                # flatten + host->device + index construction (copy stream)
        stream.wait_event(e_img)
        if prev_out[0] is not None:
            stream.wait_event(prev_out[0])   # the previous step's results left the device before they are overwritten
        ext.extract_device(d_in[i % 2], d_kps, d_desc, d_cnt, stream=stream)
        mat.bruteforce_sets_device(d_desc, d_cnt, q_set, t_set, out=d_match, stream=stream)
        e_em = stream.record_event()
        o.run()                              # compute stream, after this handle's upload
        with torch.cuda.stream(out_stream):
            out_stream.wait_event(e_em)
            h_cnt.copy_(d_cnt, non_blocking=True)
            h_kps.copy_(d_kps, non_blocking=True)
            h_desc.copy_(d_desc, non_blocking=True)
            for hm, dm in zip(h_match, d_match):
                hm.copy_(dm, non_blocking=True)
            prev_out[0] = out_stream.record_event()
        r = 0
        if not first:
            opts[(i - 1) % 2].download_batch(out=h_ba)     # results of the previous step's windows (waits for that run only)
            r = int(h_ba[2].sum())
        prev_out[0].synchronize()
        return int(h_cnt.sum()) + r          # the step's results are read on the host

    def e2e_run(n):
        prev_out[0] = None
        for i in range(n):
            e2e_step(i, i == 0)
        opts[(n - 1) % 2].download_batch(out=h_ba)
        return int(h_ba[2].sum())

    e2e_run(max(2, args.warmup // 2))
    barrier()
    opt2.synchronize()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    opt2.synchronize()
    e2e_s = time.perf_counter() - t0
    ba_in = sum(sum(np.asarray(p[k]).nbytes for k in ("poses", "pose_fixed", "points", "edge_pose", "edge_point", "edge_cam", "edge_obs",
                                                        "edge_inv_sigma2", "cam_K", "cam_ext", "cam_adj")) for p in ba_problems)
    h2d = host[0].numel() + ba_in
    d2h = sum(t.numel() * t.element_size() for t in (h_cnt, h_kps, h_desc) + h_match + h_ba)

    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    if rank == 0:
        peak, peak_kind = peaks()
        total_frames = F * world
        fps = total_frames * args.steps / (ms / 1e3)
        e2e_fps = total_frames * args.steps / (e2e_ms / 1e3)
        NI = F * CAMS
        nc = max(calls, 1)
        # per-kernel roofline table: (ms per bench step, algorithmic bytes per bench step)
        kern = {
            "resize_level_kernel(x7)": (stage_ms["pyramid"] / nc, (PYR_PIXELS_1UP + 926546) * NI),
            "fast_cells_kernel": (stage_ms["fast"] / nc, PYR_PIXELS * NI),
            "quadtree_kernel": (stage_ms["quadtree"] / nc, None),
            "describe_kernel": (stage_ms["describe"] / nc, (1849 + 60) * n_kp),
            "bruteforce_kernel": (match_ms / max(mcalls, 1), BYTES_MATCH_PAIR * NI),
        }
        bb = ba_bytes(ba_problems, ba_stats)
        for k, v in (ba_kernel_ms or {}).items():
            kern["ba_" + k] = (v / args.steps, bb[k])
        table = {}
        for k, (kms, kb) in kern.items():
            table[k] = {"ms_per_step": kms, "algorithmic_bytes_per_step": kb,
                        "achieved_gbs": (kb / (kms * 1e-3) / 1e9 if kb and kms > 0 else None)}
            if table[k]["achieved_gbs"] is not None:
                table[k]["frac"] = table[k]["achieved_gbs"] / peak
        dom_name = max((k for k in kern if kern[k][1]), key=lambda k: kern[k][0])
        dom_ms, dom_bytes = kern[dom_name]
        n_dom_launch = (ba_steps / args.steps) if dom_name.startswith("ba_") else (7 if dom_name.startswith("resize") else 1)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        step_bytes = (BYTES_EXTRACT_IMAGE * CAMS + BYTES_MATCH_PAIR * CAMS) * F + sum(bb.values())
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom_name)
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": fps, "unit": "dual-frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (extract, match) / f64 (LocalBA)",
            "data": "synthetic",
            "config": {"workload": workload_text(F, KF), "frames_per_step_per_gpu": F, "ba_windows_per_step_per_gpu": NBA, "kf_interval": KF,
                       "stages": STAGES, "keypoints_per_step": n_kp,
                       "ba": {"edges_per_window": nEt // max(NBA, 1), "lm_iterations": ba_stats[0]["iterations"], "lm_trials": ba_stats[0]["trials"],
                              "distinct_windows": min(BA_UNIQUE, NBA)},
                       "l2": "two alternating input batches of 157 MB each (> 126 MB L2); BA working set 17 MB per window",
                       "parallelism": f"{world} independent replicas, no collective"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_fps, "unit": "dual-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_kind, "traffic": traffic, "algorithmic_bytes_per_launch": dom_bytes / max(n_dom_launch, 1),
                         "ms_per_launch": dom_ms / max(n_dom_launch, 1), "launches_per_step": n_dom_launch,
                         "step_frac_of_hbm_roofline": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "stage_ms_per_step": {**{k: v / nc for k, v in stage_ms.items()}, "match": match_ms / max(mcalls, 1),
                                               "localBA": ba_ms / args.steps},
                         "kernels": table},
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample
            t = cpu_path(a[:sample], ba_problems[:BA_UNIQUE], KF, 1)
            out["cpu_baseline"] = {"value": sample / t, "unit": "dual-frames/s", "cores": 1, "kind": "port",
                                   "sample": f"first {sample} dual-frames of the same batch, single thread, {STAGES} ({t:.1f} s)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--kf-interval", type=int, default=1, help="dual-frames per LocalBA window")
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        # the ranks of a node share its host cores: split them for the library's host-side flattening of the BA graphs
        os.environ.setdefault("ORB_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // max(world, 1))))
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
