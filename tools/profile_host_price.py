#!/usr/bin/env python
"""GPU experiment: host-side cost of MonteCarloTechnique.price() -- cProfile over many pricings of a
job small enough that the device time is negligible next to the Python / ctypes / torch overhead."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob

call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
stock, rate, model = ob.Stock(spot=100.0), ob.Rate(rate=0.05), ob.HestonModel()
tech = ob.MonteCarloTechnique(n_paths=1 << 14, n_steps=16, seed=42)
for _ in range(50):
    tech.price(call, stock, model, rate)
torch.cuda.synchronize()
n = 3000
t0 = time.perf_counter()
for _ in range(n):
    tech.price(call, stock, model, rate)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / n * 1e6:.1f} us per price() (2^14 paths x 16 steps)")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    tech.price(call, stock, model, rate)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
