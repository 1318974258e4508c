"""kernel time of LocalBA as a function of the lock-step batch size (L2 residency of the edge records)"""
import os, sys, json, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from orbslam2_dualcam_b200 import Optimizer
probs = bench.make_ba(0, 256)
dev = torch.device('cuda:0')
res = {}
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
for nb in (256, 128, 64, 32):
    opts = []
    for i in range(0, 256, nb):
        o = Optimizer(max_problems=nb, device=0)
        o.upload(o.prepare(probs[i:i + nb]))
        o.profile(True)
        o.set_stream(stream)
        opts.append(o)
    torch.cuda.synchronize()
    tot = None; wall = 0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for o in opts: o.run()
        sm = sum(o.stage_ms()[0] for o in opts)
        e1.record(); torch.cuda.synchronize()
        wall = e0.elapsed_time(e1)
        t = None
        for o in opts:
            km, ks = o.kernel_ms()
            t = km if t is None else {k: t[k] + v for k, v in km.items()}
        tot = t
    res[nb] = dict(wall_ms=wall, stage_ms=sm, kernels=tot, sum=sum(tot.values()))
    print(nb, json.dumps(res[nb]), flush=True)
    del opts
