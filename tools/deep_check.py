#!/usr/bin/env python
"""GPU check at high statistical resolution: with n_steps = 1 the log-Euler BSM, Merton and Kou
schemes sample the exact terminal law, so the Monte Carlo price must agree with the reference's
closed-form / FFT price (tests/golden/price_goldens.json) to within its standard error -- here
with 2^32 paths per model, i.e. standard errors of ~1e-4 on a price of ~10.  A bias in the
device normals, the Poisson draw or the jump sizes at the 1e-5 level would show as |z| >> 3."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optpricing_b200 as ob

g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "price_goldens.json")))
call = ob.Option(100.0, 1.0, ob.OptionType.CALL)
stock, rate = ob.Stock(100.0), ob.Rate(0.05)
n_chunks = int(os.environ.get("CHUNKS", "16"))
for name, model, want in (("bsm", ob.BSMModel({"sigma": 0.2}), g["G1_closed_form"]),
                          ("merton", ob.MertonJumpModel(), g["merton_closed_form"]),
                          ("kou", ob.KouModel(), g["kou_fft"])):
    for precision in ("fp64", "fp32"):
        t = ob.MonteCarloTechnique(n_paths=1 << 28, n_steps=1, seed=2026, precision=precision)
        s = ss = 0.0
        for _ in range(n_chunks):
            p = t.price(call, stock, model, rate).price
            s += p
            ss += t.last_stats["stderr"] ** 2
        price, se = s / n_chunks, math.sqrt(ss) / n_chunks
        print(f"{name:7s} {precision}: {n_chunks} x 2^28 paths  price {price:.6f} +- {se:.6f}  "
              f"exact {want:.6f}  z = {(price - want) / se:+.2f}", flush=True)
