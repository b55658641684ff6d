#!/usr/bin/env python
"""GPU experiment: AmericanMonteCarloTechnique.price for all seven SDE models at BASELINE
config 4's size (2^22 paths x 50 exercise dates, cubic basis, put K=110)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob

aput = ob.Option(110.0, 1.0, ob.OptionType.PUT)
stock, rate = ob.Stock(100.0, dividend=0.01), ob.Rate(0.05)
n_paths, n_steps = 1 << 22, 50
for name, model in (("bsm", ob.BSMModel()), ("heston", ob.HestonModel()), ("merton", ob.MertonJumpModel()),
                    ("bates", ob.BatesModel()), ("kou", ob.KouModel()), ("sabr", ob.SABRModel()),
                    ("sabr_jump", ob.SABRJumpModel())):
    t = ob.AmericanMonteCarloTechnique(n_paths=n_paths, n_steps=n_steps, seed=1, lsm_degree=3)
    for _ in range(2):
        p = t.price(aput, stock, model, rate).price
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        p = t.price(aput, stock, model, rate).price
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 200
    print(f"american put {name:9s} 2^22 x 50 deg 3: {ms:7.3f} ms per price = "
          f"{n_paths * n_steps / ms / 1e6:6.1f} G path-steps/s  price {p:.4f}", flush=True)
