"""GlobalBundleAdjustemnt over the GPUs of a node (BASELINE configs[4]): one process per GPU, landmark-partitioned, NCCL all-reduce of
the reduced camera system.  Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gba_dist_run.py [n_kf n_points its]
(or plain `python tools/gba_dist_run.py` for one GPU).  Rank 0 prints one JSON line; with --check the result is compared with a single-GPU run."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from orbslam2_dualcam_b200 import DistributedOptimizer, shard_problem
import synth


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_kf, n_pts, its = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (2000, 200000, 5)
    check = "--check" in sys.argv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = synth.gba_problem(0, n_kf=n_kf, n_points=n_pts)
    opt = DistributedOptimizer.from_torch_distributed(local) if world > 1 else DistributedOptimizer(device=local)
    sh = shard_problem(p, rank, world)
    opt.GlobalBundleAdjustemnt(sh, nIterations=1)                      # warm-up (allocations, NCCL channels)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    poses, points, st = opt.GlobalBundleAdjustemnt(sh, nIterations=its)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    tm = opt.timing()
    ok = None
    if check:
        if rank == 0:
            ref = DistributedOptimizer(device=local)
            rposes, _, rst = ref.GlobalBundleAdjustemnt(shard_problem(p, 0, 1), nIterations=its)
            rel = float((np.linalg.norm(poses - rposes, axis=1) / np.linalg.norm(rposes, axis=1)).max())
            ok = {"pose_rel_vs_1gpu": rel, "trials_equal": rst["trials"] == st["trials"]}
    if rank == 0:
        n = 6 * int((p["pose_fixed"] == 0).sum())
        print(json.dumps({"what": "GlobalBundleAdjustemnt", "n_gpus": world, "key_frames": n_kf, "points": n_pts, "edges": int(len(p["edge_pose"])),
                          "lm_iterations": st["iterations"], "lm_trials": st["trials"], "seconds": dt.item(), "ms_per_trial": tm["loop_ms"] / max(st["trials"], 1),
                          "ms_per_trial_wall_incl_host_setup": 1e3 * dt.item() / max(st["trials"], 1),
                          "reduced_system": n, "skyline_blocks": tm["skyline_blocks"], "skyline_MB": tm["skyline_blocks"] * 288 / 1e6,
                          "allreduce_ms_per_trial": tm["allreduce_ms"] / max(st["trials"], 1),
                          "solve_ms_per_trial": tm["solve_ms"] / max(st["trials"], 1),
                          "allreduce_GBps": (tm["allreduce_bytes"] / 1e9) / max(tm["allreduce_ms"] / 1e3, 1e-9) if world > 1 else None,
                          "chi2": [st["initial_chi2"], st["final_chi2"]], "check": ok}))
    if world > 1:
        dist.destroy_process_group()


main()
