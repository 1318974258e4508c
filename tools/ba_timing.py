"""Device time of the batched LocalBA kernel: N copies of the config-3 problem (one persistent CTA each)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from orbslam2_dualcam_b200 import Optimizer
import synth

def main():
    ns = [int(x) for x in sys.argv[1:]] or [1, 32, 148, 296]
    base = [synth.ba_problem(s) for s in range(4)]
    for n in ns:
        opt = Optimizer(max_problems=n)
        t0 = time.perf_counter()
        opt.upload([base[i % 4] for i in range(n)])
        t_up = time.perf_counter() - t0
        opt.profile(True)
        opt.run(); opt.synchronize()
        ms = []
        for _ in range(3):
            opt.run()
            m, c = opt.stage_ms()
            ms.append(m)
        st = opt.download(0)[3]
        print(f"n={n} upload {t_up*1e3:.1f} ms  kernel {min(ms):.2f} ms  -> {n/min(ms)*1e3:.1f} problems/s  its={st['iterations']} trials={st['trials']}", flush=True)
        opt.close()
main()
