"""Host- and link-side cost of the pieces of the end-to-end step, per rank, with all ranks of the node active at once
(python tools/e2e_pieces.py, or under torchrun for N GPUs).  Prints one line per rank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("ORB_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // max(world, 1))))
from orbslam2_dualcam_b200 import Optimizer, compact_problem
import synth
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bar = (lambda: dist.barrier()) if world > 1 else (lambda: None)
n = 256
base = [synth.ba_problem(s) for s in range(4)]
lev = synth.inv_sigma2_levels()
cbase = [compact_problem(p, lev) for p in base]
opt = Optimizer(max_problems=n, device=local)
prep = Optimizer.prepare_f32([cbase[i % 4] for i in range(n)])
prep64 = opt.prepare([base[i % 4] for i in range(n)])
h = torch.empty((256, 2, 480, 640), dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
res = {}
for k in range(3):
    bar(); torch.cuda.synchronize()
    t0 = time.perf_counter(); opt.upload(prep); t1 = time.perf_counter(); opt.synchronize(); t2 = time.perf_counter()
    res["upload_f32_host_ms"], res["upload_f32_device_tail_ms"] = 1e3 * (t1 - t0), 1e3 * (t2 - t1)
    bar(); torch.cuda.synchronize()
    t0 = time.perf_counter(); opt.upload(prep64); t1 = time.perf_counter(); opt.synchronize(); t2 = time.perf_counter()
    res["upload_f64_host_ms"], res["upload_f64_device_tail_ms"] = 1e3 * (t1 - t0), 1e3 * (t2 - t1)
    bar(); torch.cuda.synchronize()
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); res["h2d_157MB_ms"] = 1e3 * (time.perf_counter() - t0)
    bar(); torch.cuda.synchronize()
    t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); res["d2h_157MB_ms"] = 1e3 * (time.perf_counter() - t0)
opt.upload(prep); opt.run(); opt.synchronize()
bar(); t0 = time.perf_counter(); out = opt.download_batch(); res["download_batch_ms"] = 1e3 * (time.perf_counter() - t0)
print(f"rank {rank}/{world} threads {os.environ['ORB_HOST_THREADS']} cpus {os.cpu_count()}: " + ", ".join(f"{k} {v:.1f}" for k, v in res.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
