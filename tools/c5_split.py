#!/usr/bin/env python
"""GPU experiment: where the C5 step (price + delta + vega, SABR-jump and Kou, 2^24 x 365,
control variate) spends its time, per model and per kernel launch of price_with_greeks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob
from optpricing_b200 import _engine

eng = _engine.get_engine()
call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
stock, rate = ob.Stock(spot=100.0), ob.Rate(rate=0.05)
for m in (ob.SABRJumpModel(), ob.KouModel(), ob.BSMModel(), ob.MertonJumpModel()):
    t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=365, seed=11, control_variate=True)
    for what, f in (("price", lambda: t.price(call, stock, m, rate).price),
                    ("price_with_greeks(delta, vega)",
                     lambda: t.price_with_greeks(call, stock, m, rate, greeks=("delta", "vega")).price)):
        best = 1e9
        for rep in range(4):
            l0 = eng.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            if rep:
                best = min(best, a.elapsed_time(b))
        print(f"{type(m).__name__:16s} {what:32s} {best:8.3f} ms, {eng.launch_count() - l0} launches", flush=True)
