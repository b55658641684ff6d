#!/usr/bin/env python
"""GPU experiment: Heston at the shard sizes of 1/2/4/8-GPU strong scaling, 768-thread blocks
against the host's choice between 768 and 1024 (euro_threads_auto)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optpricing_b200 import _engine
eng = _engine.get_engine()
H = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5}
mdl = _engine.make_model("heston", H, 0.05, 0.0)
mom = torch.zeros(6, dtype=torch.float64, device=eng.device)
for lg in (21, 22, 23, 24):
    for auto in (0, 1):
        eng.set_int("euro_threads_auto", auto)
        best = 1e9
        for rep in range(6):
            sim = eng.make_sim(100.0, 1.0, (1 << lg) // 2, 252, True, 100, 0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.european_async(mdl, sim, [100.0], [1], mom); b.record(); torch.cuda.synchronize()
            if rep: best = min(best, a.elapsed_time(b))
        print(f"2^{lg} paths auto={auto}: {best:.3f} ms  ({(1<<lg)*252/best/1e6:.1f} G/s)  price {float(mom[1]/mom[0])*0.951229:.6f}", flush=True)
