#!/usr/bin/env python
"""GPU experiment: BASELINE config 4 -- American put, GBM, 2^22 paths x 50 exercise
dates, cubic basis.  Times the Philox path-store fill and the Longstaff-Schwartz
induction separately (CUDA events) and reports HBM GB/s against the algorithmic
32 B per path-date."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob
from optpricing_b200 import _engine

eng = _engine.get_engine()
eng.set_int("lsmc_mode", int(os.environ.get("LSMC_MODE", "1")))
n_paths = int(os.environ.get("N_PATHS", 1 << 22)); n_steps = int(os.environ.get("N_STEPS", 50))
degree = int(os.environ.get("DEGREE", 3))
mdl = _engine.make_model("bsm", {"sigma": 0.2}, 0.05, 0.01)
store, ld = eng.lsmc_store(n_paths, n_steps)
ev = lambda: torch.cuda.Event(enable_timing=True)
best_fill = best_ind = 1e9
for rep in range(4):
    sim = eng.make_sim(100.0, 1.0, n_paths // 2, n_steps, True, 7 + rep, 1)
    a, b, c = ev(), ev(), ev()
    a.record()
    eng.lsmc_fill_paths(mdl, sim, store, ld)
    b.record()
    out = eng.lsmc(store, ld, n_paths, n_steps, 110.0, False, 0.05, 1.0 / n_steps, degree)
    c.record()
    torch.cuda.synchronize()
    if rep:
        best_fill = min(best_fill, a.elapsed_time(b)); best_ind = min(best_ind, b.elapsed_time(c))
price = out[1] / out[0]
se = ((out[2] / out[0] - price ** 2) / out[0]) ** 0.5
pd = n_paths * n_steps
print(f"[mode {os.environ.get('LSMC_MODE', '1')}] LSMC {n_paths} x {n_steps} deg {degree}: fill {best_fill:.3f} ms ({pd * 8 / best_fill / 1e6:.0f} GB/s store, "
      f"{pd / best_fill / 1e6:.1f} G path-steps/s), induction {best_ind:.3f} ms "
      f"({pd * 32 / best_ind / 1e6:.0f} GB/s at 32 B/path-date), total {best_fill + best_ind:.3f} ms = "
      f"{pd / (best_fill + best_ind) / 1e6:.2f} G path-steps/s; price {price:.4f} +- {se:.4f} (CRR 12.2777)")
t = ob.AmericanMonteCarloTechnique(n_paths=n_paths, n_steps=n_steps, seed=1, lsm_degree=degree)
opt = ob.Option(110.0, 1.0, ob.OptionType.PUT); st = ob.Stock(100.0, dividend=0.01); rt = ob.Rate(0.05)
t.price(opt, st, ob.BSMModel(), rt)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    p = t.price(opt, st, ob.BSMModel(), rt).price
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"technique.price() end to end: {dt * 1e3:.3f} ms = {pd / dt / 1e9:.2f} G path-steps/s; price {p:.4f}")
