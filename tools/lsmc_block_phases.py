#!/usr/bin/env python
"""GPU experiment: per-BLOCK sweep and wait times of k_lsmc_persistent (every block stamps
clock64 at four points per date): which blocks are the slow ones, and by how much."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optpricing_b200 import _engine
eng = _engine.get_engine()
if "LSMC_RING" in os.environ:
    eng.set_int("lsmc_ring", int(os.environ["LSMC_RING"]))
n_steps, degree, G = 50, 3, torch.cuda.get_device_properties(0).multi_processor_count
mdl = _engine.make_model("bsm", {"sigma": 0.2}, 0.05, 0.01)
for n_paths in [int(x) for x in os.environ.get("SIZES", "2097152,4194304").split(",")]:
    store, ld = eng.lsmc_store(n_paths, n_steps)
    sim = eng.make_sim(100.0, 1.0, n_paths // 2, n_steps, True, 7, 1)
    eng.lsmc_fill_paths(mdl, sim, store, ld)
    tim = torch.zeros(G * (n_steps + 1) * 4, dtype=torch.int64, device=eng.device)
    eng.set_int("lsmc_timing_all", 1)
    for rep in range(3):
        eng.set_ptr("lsmc_timing", tim if rep == 2 else None)
        eng.lsmc(store, ld, n_paths, n_steps, 110.0, False, 0.05, 1.0 / n_steps, degree)
    eng.set_ptr("lsmc_timing", None)
    eng.set_int("lsmc_timing_all", 0)
    t = tim.cpu().numpy().reshape(G, n_steps + 1, 4).astype(np.float64) / 1965.0
    rows = t[:, 1:n_steps][:, ::-1]                      # [block][date][stamp]
    sweep = np.median(rows[:, :, 3] - rows[:, :, 2], axis=1)
    wait = np.median(rows[:, :, 1] - rows[:, :, 0], axis=1)
    order = np.argsort(sweep)
    print(f"n_paths {n_paths}: sweep per block [us] min {sweep.min():.2f} median {np.median(sweep):.2f} "
          f"p90 {np.percentile(sweep, 90):.2f} max {sweep.max():.2f};  wait median {np.median(wait):.2f} min {wait.min():.2f}")
    print("  slowest blocks:", [(int(b), round(float(sweep[b]), 2)) for b in order[-12:]])
    print("  fastest blocks:", [(int(b), round(float(sweep[b]), 2)) for b in order[:8]])
    print("  sweep by block:", " ".join(f"{x:.1f}" for x in sweep))
