"""Host-side cost of the pieces of the end-to-end step (second call of each, after allocations)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from orbslam2_dualcam_b200 import Optimizer, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = [synth.ba_problem(s) for s in range(4)]
probs = [base[i % 4] for i in range(n)]
opt = Optimizer(max_problems=n)
prep = opt.prepare(probs)
for k in range(3):
    t0 = time.perf_counter(); opt.upload(prep); t1 = time.perf_counter(); opt.synchronize(); t2 = time.perf_counter()
    opt.run(); opt.synchronize(); t3 = time.perf_counter()
    out = opt.download_batch(); t4 = time.perf_counter()
    print(f"upload call {1e3*(t1-t0):.1f} ms (+{1e3*(t2-t1):.1f} ms to finish on device), run {1e3*(t3-t2):.1f} ms, download_batch {1e3*(t4-t3):.1f} ms", flush=True)
nP = sum(len(p["pose_fixed"]) for p in probs); nL = sum(len(p["points"]) for p in probs); nE = sum(len(p["edge_pose"]) for p in probs)
pin = (torch.empty((nP, 12), dtype=torch.float64).pin_memory(), torch.empty((nL, 3), dtype=torch.float64).pin_memory(), torch.empty((nE,), dtype=torch.uint8).pin_memory())
t0 = time.perf_counter(); opt.download_batch(out=pin); print(f"download_batch pinned {1e3*(time.perf_counter()-t0):.1f} ms")
h = torch.empty((256, 2, 480, 640), dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print(f"H2D 157 MB pinned {1e3*(time.perf_counter()-t0):.1f} ms")
