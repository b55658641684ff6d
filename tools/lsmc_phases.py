#!/usr/bin/env python
"""GPU experiment: where one exercise date of k_lsmc_persistent spends its time.
Block 0 stamps clock64 at four points per date (see lsmc.cuh); prints the median
duration of each phase over the dates (SM cycles at 1.965 GHz -> microseconds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optpricing_b200 import _engine
eng = _engine.get_engine()
n_steps, degree = 50, int(os.environ.get("DEGREE", 3))
if "LSMC_RING" in os.environ:
    eng.set_int("lsmc_ring", int(os.environ["LSMC_RING"]))
mdl = _engine.make_model("bsm", {"sigma": 0.2}, 0.05, 0.01)
for n_paths in [int(x) for x in os.environ.get("SIZES", "262144,1048576,2097152,4194304").split(",")]:
    store, ld = eng.lsmc_store(n_paths, n_steps)
    sim = eng.make_sim(100.0, 1.0, n_paths // 2, n_steps, True, 7, 1)
    eng.lsmc_fill_paths(mdl, sim, store, ld)
    tim = torch.zeros((n_steps + 1) * 4, dtype=torch.int64, device=eng.device)
    for rep in range(3):
        eng.set_ptr("lsmc_timing", tim if rep == 2 else None)
        eng.lsmc(store, ld, n_paths, n_steps, 110.0, False, 0.05, 1.0 / n_steps, degree)
    eng.set_ptr("lsmc_timing", None)
    t = tim.cpu().numpy().reshape(n_steps + 1, 4).astype(np.float64)
    # row d+1 holds the stamps of the iteration that sweeps date d; iterations run d = n-1 .. 0
    rows = t[1:n_steps][::-1]            # dates n-2 .. 0 (skip the first: no exchange)
    poll = rows[:, 1] - rows[:, 0]
    solve = rows[:, 2] - rows[:, 1]
    sweep = rows[:, 3] - rows[:, 2]
    nxt = np.r_[rows[1:, 0], t[0, 0]]
    publish = nxt - rows[:, 3]
    med = lambda a: float(np.median(a)) / 1965.0
    print(f"n_paths {n_paths}: per date [us] poll+sum {med(poll):.2f}  reduce+solve {med(solve):.2f}  "
          f"sweep {med(sweep):.2f}  block_sum+publish {med(publish):.2f}  total {med(poll+solve+sweep+publish):.2f}")
