#!/usr/bin/env python
"""GPU experiment: compact (k_european, several blocks per SM) against lane-replicated
(k_european_wide, one block per SM) tables, per model kind, 2^24 x 252 by default.
Library variants through OPTMC_LIB.  Prints ms per pricing and the price of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optpricing_b200 import _engine

eng = _engine.get_engine()
HESTON = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5}
PARAMS = {"heston": HESTON, "bsm": {"sigma": 0.2},
          "bates": {**HESTON, "lambda": 0.5, "mu_j": -0.1, "sigma_j": 0.15},
          "merton": {"sigma": 0.2, "lambda": 0.5, "mu_j": -0.1, "sigma_j": 0.15},
          "kou": {"sigma": 0.15, "lambda": 1.0, "p_up": 0.6, "eta1": 10.0, "eta2": 5.0},
          "sabr": {"alpha": 0.5, "beta": 0.8, "rho": -0.6},
          "sabr_jump": {"alpha": 0.5, "beta": 0.8, "rho": -0.6, "lambda": 0.4, "mu_j": -0.1, "sigma_j": 0.15}}
kinds = sys.argv[1:] or ["heston", "bsm", "bates", "merton", "kou"]
n_paths, n_steps = int(os.environ.get("N_PATHS", 1 << 24)), int(os.environ.get("N_STEPS", 252))
mom = torch.zeros(6, dtype=torch.float64, device=eng.device)
for kind in kinds:
    mdl = _engine.make_model(kind, PARAMS[kind], 0.05, 0.0)
    for wide in (0, 1):
        eng.set_int("euro_wide", wide)
        best = 1e9
        for rep in range(5):
            sim = eng.make_sim(100.0, 1.0, n_paths // 2, n_steps, True, 100, 0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.european_async(mdl, sim, [100.0], [1], mom)
            b.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, a.elapsed_time(b))
        m = mom.cpu().numpy()
        se = ((m[2] / m[0] - (m[1] / m[0]) ** 2) / m[0]) ** 0.5 * 0.951229
        print(f"{os.path.basename(_engine.LIB_PATH):22s} {kind:7s} wide={wide}: {n_paths}x{n_steps} "
              f"{best:7.3f} ms  {n_paths * n_steps / best / 1e6:8.1f} G path-steps/s  "
              f"price {m[1] / m[0] * 0.951229:.5f} +- {se:.5f}", flush=True)
