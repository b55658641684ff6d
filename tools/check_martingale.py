#!/usr/bin/env python
"""GPU check: z-scores of the sample mean of S_T against the scheme's exact
martingale mean, over several seeds.  A biased RNG / step shows up as a
non-zero mean z."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optpricing_b200 as ob
from optpricing_b200 import _engine
eng = _engine.get_engine()
impl = int(os.environ.get("IMPL", "1"))
eng.set_int("normals_impl", impl)
call = ob.Option(100.0, 1.0, ob.OptionType.CALL)
stock, rate = ob.Stock(100.0), ob.Rate(0.05)
for name, model, corr in (("bsm", ob.BSMModel(), "reference"), ("heston", ob.HestonModel(), "reference"),
                          ("heston-indep", ob.HestonModel(), "independent"), ("kou", ob.KouModel(), "reference")):
    zs, zp = [], []
    for seed in range(int(os.environ.get("NSEEDS", "12"))):
        t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=252, seed=1000 + seed, correlation=corr)
        t.price(call, stock, model, rate)
        s = t.last_stats
        zs.append((s["control_mean"] - s["control_expected"]) / s["control_stderr"])
    zs = np.array(zs)
    print(f"impl={impl} {name:13s} z: mean {zs.mean():+.2f} (se {1/math.sqrt(len(zs)):.2f}) min {zs.min():+.2f} max {zs.max():+.2f}  "
          + " ".join(f"{z:+.1f}" for z in zs), flush=True)
