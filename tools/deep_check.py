#!/usr/bin/env python
"""GPU check at high statistical resolution: with n_steps = 1 the log-Euler BSM, Merton and Kou
schemes sample the exact terminal law, so the Monte Carlo price must agree with the reference's
closed-form price (tests/golden/price_goldens.json; Kou: Gil-Pelaez quadrature) to within its standard error -- here
with 2^32 paths per model, i.e. standard errors of ~1e-4 on a price of ~10.  A bias in the
device normals, the Poisson draw or the jump sizes at the 1e-5 level would show as |z| >> 3.
Heston with vol_of_vol -> 0 (1e-9 here: the model wants it positive) and v0 = theta has a constant variance, so its log-Euler scheme is exact
for ANY number of steps and must reproduce the Black-Scholes price with sigma = sqrt(v0): that
holds the two-factor step -- the mixed Box-Muller direction and the one-root sqrt(v+ * (-2 ln u))
-- to the same resolution (8 steps, both correlation conventions)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optpricing_b200 as ob

def kou_exact(S0, K, T, r, q, p):
    """Kou call by Gil-Pelaez quadrature of the characteristic function (the reference's FFT
    price for the same contract, 11.567506, is off by 5.6e-4: its grid error)."""
    import numpy as np
    from scipy.integrate import quad
    sg, lam, pu, e1, e2 = p["sigma"], p["lambda"], p["p_up"], p["eta1"], p["eta2"]
    comp = lam * ((pu * e1 / (e1 - 1)) + ((1 - pu) * e2 / (e2 + 1)) - 1)

    def cf(u):
        drift = math.log(S0) + (r - q - 0.5 * sg ** 2 - comp) * T
        return np.exp(1j * u * drift - 0.5 * sg ** 2 * u ** 2 * T
                      + lam * T * (pu * e1 / (e1 - 1j * u) + (1 - pu) * e2 / (e2 + 1j * u) - 1))
    k = math.log(K)
    f1 = lambda u: np.real(np.exp(-1j * u * k) * cf(u - 1j) / (1j * u * cf(-1j)))   # noqa: E731
    f2 = lambda u: np.real(np.exp(-1j * u * k) * cf(u) / (1j * u))                  # noqa: E731
    P1 = 0.5 + quad(f1, 1e-12, 400, limit=2000, epsabs=1e-13, epsrel=1e-12)[0] / math.pi
    P2 = 0.5 + quad(f2, 1e-12, 400, limit=2000, epsabs=1e-13, epsrel=1e-12)[0] / math.pi
    return math.exp(-r * T) * (float(np.real(cf(-1j))) * P1 - K * P2)


g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "price_goldens.json")))
call = ob.Option(100.0, 1.0, ob.OptionType.CALL)
stock, rate = ob.Stock(100.0), ob.Rate(0.05)
n_chunks = int(os.environ.get("CHUNKS", "16"))
for name, model, want in (("bsm", ob.BSMModel({"sigma": 0.2}), g["G1_closed_form"]),
                          ("merton", ob.MertonJumpModel(), g["merton_closed_form"]),
                          ("kou", ob.KouModel(), kou_exact(100.0, 100.0, 1.0, 0.05, 0.0, ob.KouModel().params))):
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

hes = ob.HestonModel({"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 1e-9})
for corr in ("reference", "independent"):
    for precision in ("fp64", "fp32"):
        t = ob.MonteCarloTechnique(n_paths=1 << 28, n_steps=8, seed=77, precision=precision,
                                   correlation=corr)
        s = ss = 0.0
        for _ in range(n_chunks):
            s += t.price(call, stock, hes, rate).price
            ss += t.last_stats["stderr"] ** 2
        price, se = s / n_chunks, math.sqrt(ss) / n_chunks
        print(f"heston xi=0 ({corr[:5]}) {precision}: {n_chunks} x 2^28 paths x 8 steps  price {price:.6f} "
              f"+- {se:.6f}  exact {g['G1_closed_form']:.6f}  z = {(price - g['G1_closed_form']) / se:+.2f}",
              flush=True)
