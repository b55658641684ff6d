#!/usr/bin/env python
"""
Time the UNMODIFIED reference (Python + numba, single thread: its kernels have no
parallel=True) on this host, from baseline/_ref (tools/populate_baseline_ref.py).

    python tools/time_reference_numba.py c2|c1|c4 [log2_paths]

Prints one JSON object: path-steps/s of `technique.price()`, warm (one untimed call first:
numba JIT), best of `REPS` (default 2).  bench.py runs this in a subprocess with a time limit
and reports it as cpu_baseline.reference_numba next to the oracle port's figures.
C2 cannot run at 2^24 paths on the reference (three 16.9 GB increment arrays):
2^20 paths by default, the rate is per path-step.
"""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "baseline", "_ref", "src", "optpricing")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else {"c1": 17, "c2": 20, "c4": 20}[name]
    reps = int(os.environ.get("REPS", "2"))
    if not os.path.isdir(SRC):
        print(json.dumps({"unavailable": "baseline/_ref not populated"}))
        return
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    shim = types.ModuleType("optpricing")           # skip the top-level __init__
    shim.__path__ = [SRC]
    sys.modules["optpricing"] = shim
    import numba
    from optpricing.atoms import Option, OptionType, Rate, Stock
    from optpricing.models import BSMModel, HestonModel
    from optpricing.techniques import AmericanMonteCarloTechnique, MonteCarloTechnique
    n = 1 << lg
    if name == "c4":
        tech = AmericanMonteCarloTechnique(n_paths=n, n_steps=50, seed=42, lsm_degree=3)
        args = (Option(strike=110.0, maturity=1.0, option_type=OptionType.PUT),
                Stock(spot=100.0, dividend=0.01), BSMModel(params={"sigma": 0.2}), Rate(rate=0.05))
        steps, label = 50, "AmericanMonteCarloTechnique.price, GBM put K=110, cubic basis"
    elif name == "c1":
        n = 100_000
        tech = MonteCarloTechnique(n_paths=n, n_steps=252, seed=42)
        args = (Option(strike=100.0, maturity=1.0, option_type=OptionType.CALL), Stock(spot=100.0),
                BSMModel(params={"sigma": 0.2}), Rate(rate=0.05))
        steps, label = 252, "MonteCarloTechnique.price, BSM call"
    else:
        tech = MonteCarloTechnique(n_paths=n, n_steps=252, seed=42)
        args = (Option(strike=100.0, maturity=1.0, option_type=OptionType.CALL), Stock(spot=100.0),
                HestonModel(params=HestonModel.default_params), Rate(rate=0.05))
        steps, label = 252, "MonteCarloTechnique.price, Heston call"
    small = type(tech)(n_paths=2048, n_steps=steps, seed=1) if name != "c4" else \
        AmericanMonteCarloTechnique(n_paths=2048, n_steps=50, seed=1, lsm_degree=3)
    small.price(*args)                                    # JIT warm-up
    best, price = float("inf"), None
    for _ in range(reps):
        t0 = time.perf_counter()
        price = float(tech.price(*args).price)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"value": n * steps / best, "unit": "path-steps/s", "cores": 1,
                      "kind": "reference", "seconds": best, "price": price,
                      "sample": f"{n} paths x {steps} steps, {label}, warm, best of {reps}; "
                                f"numba {numba.__version__}, {numba.get_num_threads()} numba threads "
                                f"available, nproc {os.cpu_count()}"}))


if __name__ == "__main__":
    main()
