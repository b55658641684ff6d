"""
Generate the golden fixtures in this directory by RUNNING THE REFERENCE.

Runs only in the build container, where /root/reference exists:

    NUMBA_CACHE_DIR=/tmp/numba_cache PYTHONDONTWRITEBYTECODE=1 \
        python tests/golden/make_golden.py

The reference cannot be imported as ``import optpricing`` (its top-level
__init__ needs package metadata and yfinance); a two-line shim
package whose __path__ points at the reference source lets the sub-packages
import untouched.  Nothing here is read at test time except the files written:

    kernel_vectors.npz   inputs + outputs of the 7 terminal and 7 path kernels
    ls_vectors.npz       a path matrix + longstaff_schwartz_pricer outputs
    price_goldens.json   technique.price()/delta/gamma/vega values
    greek_goldens.json   GreekMixin values (European BSM put, Heston call; American put)
    european_goldens.json  European puts of all seven SDE models on ragged shapes
    american_goldens.json  American prices of all seven SDE models (numpy-mode parity)
    levy_goldens.json    single-step models (VG, NIG, CGMY, Hyperbolic, CEV): seeded prices,
                         large-sample prices / quantiles, FFT prices

The GPU box has no /root/reference; tests compare the oracle and the CUDA path
against these files.
"""
from __future__ import annotations

import json
import math
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/optpricing"


def import_reference():
    shim = tempfile.mkdtemp(prefix="refshim_")
    os.makedirs(os.path.join(shim, "optpricing"))
    with open(os.path.join(shim, "optpricing", "__init__.py"), "w") as f:
        f.write(f'__path__ = ["{REF_SRC}"]\n')
    sys.path.insert(0, shim)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")


import_reference()
import numba  # noqa: E402
from optpricing.atoms import Option, OptionType, Rate, Stock  # noqa: E402
from optpricing.models import (  # noqa: E402
    BatesModel, BSMModel, HestonModel, KouModel, MertonJumpModel, SABRJumpModel, SABRModel,
)
from optpricing.techniques import (  # noqa: E402
    AmericanMonteCarloTechnique, ClosedFormTechnique, CRRTechnique, FFTTechnique,
    MonteCarloTechnique,
)
from optpricing.techniques.kernels import mc_kernels, path_kernels  # noqa: E402
from optpricing.techniques.kernels.american_mc_kernels import (  # noqa: E402
    longstaff_schwartz_pricer,
)


@numba.njit
def seed_numba(k):
    np.random.seed(k)


@numba.njit
def numba_normal(mu, sig, n):
    return np.random.normal(mu, sig, n)


@numba.njit
def numba_kou(n, p_up, eta1, eta2):
    u = np.random.random(n)
    up = np.random.exponential(1.0 / eta1, n)
    dn = -np.random.exponential(1.0 / eta2, n)
    return np.where(u < p_up, up, dn)


# ---------------------------------------------------------------- kernels
D, M = 48, 12                      # draws x steps of the kernel vectors
COMMON = dict(r=0.05, q=0.01, dt=1.0 / 50)
SCALARS = {
    "bsm": dict(log_s0=math.log(100.0), sigma=0.2),
    "heston": dict(log_s0=math.log(100.0), v0=0.04, kappa=2.0, theta=0.04, rho=-0.7,
                   vol_of_vol=1.5),          # big xi: the v=0 floor is hit often
    "merton": dict(log_s0=math.log(100.0), sigma=0.2, lambda_=15.0, mu_j=-0.1, sigma_j=0.15),
    "bates": dict(log_s0=math.log(100.0), v0=0.04, kappa=2.0, theta=0.04, rho=-0.7,
                  vol_of_vol=0.5, lambda_=15.0, mu_j=-0.1, sigma_j=0.15),
    "sabr": dict(s0=100.0, v0=0.5, alpha=0.5, beta=0.8, rho=-0.6),
    "sabr_jump": dict(s0=100.0, v0=0.5, alpha=0.5, beta=0.8, rho=-0.6, lambda_=15.0,
                      mu_j=-0.1, sigma_j=0.15),
    "kou": dict(log_s0=math.log(100.0), sigma=0.15, lambda_=15.0, p_up=0.6, eta1=10.0,
                eta2=5.0),
}
TERMINAL = dict(bsm=mc_kernels.bsm_kernel, heston=mc_kernels.heston_kernel,
                merton=mc_kernels.merton_kernel, bates=mc_kernels.bates_kernel,
                sabr=mc_kernels.sabr_kernel, sabr_jump=mc_kernels.sabr_jump_kernel,
                kou=mc_kernels.kou_kernel)
PATH = dict(bsm=path_kernels.bsm_path_kernel, heston=path_kernels.heston_path_kernel,
            merton=path_kernels.merton_path_kernel, bates=path_kernels.bates_path_kernel,
            sabr=path_kernels.sabr_path_kernel, sabr_jump=path_kernels.sabr_jump_path_kernel,
            kou=path_kernels.kou_path_kernel)
TWO_FACTOR = ("heston", "bates", "sabr", "sabr_jump")
JUMPS = ("merton", "bates", "kou", "sabr_jump")


def jump_buffer(kind, sc, counts, seed):
    """What the kernel will pull from numba's RNG after seed_numba(seed)."""
    seed_numba(seed)
    chunks = []
    for n in counts.sum(axis=0):
        if n > 0:
            if kind == "kou":
                chunks.append(numba_kou(int(n), sc["p_up"], sc["eta1"], sc["eta2"]))
            else:
                chunks.append(numba_normal(sc["mu_j"], sc["sigma_j"], int(n)))
    return np.concatenate(chunks) if chunks else np.zeros(0)


def kernel_vectors():
    out = {}
    rng = np.random.default_rng(20261019)
    for case in ("rand", "zero"):
        for kind, sc in SCALARS.items():
            kw = dict(n_paths=D, n_steps=M, **COMMON, **sc)
            scale = math.sqrt(COMMON["dt"]) if case == "rand" else 0.0
            if kind in TWO_FACTOR:
                kw["dw1"] = rng.standard_normal((D, M)) * scale
                kw["dw2"] = rng.standard_normal((D, M)) * scale
            else:
                kw["dw"] = rng.standard_normal((D, M)) * scale
            if kind in JUMPS:
                lam_dt = sc["lambda_"] * COMMON["dt"] if case == "rand" else 0.0
                kw["jump_counts"] = rng.poisson(lam_dt, (D, M))
                seed = 1000 + len(out)
                jb = jump_buffer(kind, sc, kw["jump_counts"], seed)
                # numba's stream must equal RandomState(seed): the oracle relies on it
                rs = np.random.RandomState(seed)
                chk = []
                for n in kw["jump_counts"].sum(axis=0):
                    if n > 0:
                        if kind == "kou":
                            u = rs.random_sample(int(n))
                            up = rs.exponential(1.0 / sc["eta1"], int(n))
                            dn = -rs.exponential(1.0 / sc["eta2"], int(n))
                            chk.append(np.where(u < sc["p_up"], up, dn))
                        else:
                            chk.append(rs.normal(sc["mu_j"], sc["sigma_j"], int(n)))
                chk = np.concatenate(chk) if chk else np.zeros(0)
                assert np.array_equal(jb, chk), f"numba RNG != RandomState for {kind}"
                out[f"{case}/{kind}/jump_buf"] = jb
                out[f"{case}/{kind}/jump_seed"] = np.int64(seed)
            for k, v in kw.items():
                if isinstance(v, np.ndarray):
                    out[f"{case}/{kind}/{k}"] = v
            if kind in JUMPS:
                seed_numba(seed)
            out[f"{case}/{kind}/terminal"] = TERMINAL[kind](**kw)
            if kind in JUMPS:
                seed_numba(seed)
            out[f"{case}/{kind}/paths"] = PATH[kind](**kw)
    out["meta/scalars"] = np.array(json.dumps({"common": COMMON, "models": SCALARS,
                                               "D": D, "M": M}))
    np.savez(os.path.join(HERE, "kernel_vectors.npz"), **out)
    return out


# ---------------------------------------------------------------- LS pricer
def ls_vectors():
    rng = np.random.default_rng(7)
    n, m = 2000, 12
    dt = 1.0 / m
    dw = rng.standard_normal((n, m)) * math.sqrt(dt)
    paths = path_kernels.bsm_path_kernel(n, m, math.log(100.0), 0.05, 0.01, 0.2, dt, dw)
    paths = np.concatenate([paths, path_kernels.bsm_path_kernel(
        n, m, math.log(100.0), 0.05, 0.01, 0.2, dt, -dw)])
    out = {"paths": paths, "dt": np.float64(dt), "r": np.float64(0.05)}
    cases = []
    for K, is_call in ((110.0, False), (100.0, False), (95.0, True), (60.0, False),
                       (400.0, True)):
        for degree in (1, 2, 3):
            v = longstaff_schwartz_pricer(paths, K, dt, 0.05, is_call, degree)
            cases.append(dict(K=K, is_call=is_call, degree=degree, price=float(v)))
    # one-step matrix: the induction loop body never runs (american_mc_kernels.py:53)
    v = longstaff_schwartz_pricer(np.ascontiguousarray(paths[:, :2]), 110.0, dt, 0.05, False, 2)
    cases.append(dict(K=110.0, is_call=False, degree=2, price=float(v), one_step=True))
    out["cases"] = np.array(json.dumps(cases))
    np.savez(os.path.join(HERE, "ls_vectors.npz"), **out)
    return cases


# ---------------------------------------------------------------- prices
def price_goldens():
    call = Option(strike=100, maturity=1.0, option_type=OptionType.CALL)
    put = Option(strike=100, maturity=1.0, option_type=OptionType.PUT)
    st = Stock(spot=100.0)
    stq = Stock(spot=100.0, dividend=0.01)
    rt = Rate(rate=0.05)
    g = {}
    bsm = BSMModel(params={"sigma": 0.2})
    g["G1_bsm_20000x10_seed0"] = MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0).price(call, st, bsm, rt).price
    g["G1_closed_form"] = ClosedFormTechnique().price(call, st, bsm, rt).price
    g["bsm_put_q_19999x7_seed3"] = MonteCarloTechnique(n_paths=19999, n_steps=7, seed=3).price(put, stq, bsm, rt).price
    g["bsm_noanti_5000x9_seed5"] = MonteCarloTechnique(n_paths=5000, n_steps=9, antithetic=False, seed=5).price(call, st, bsm, rt).price
    hes = HestonModel(params=HestonModel.default_params)
    g["G2_heston_20000x100_seed42"] = MonteCarloTechnique(n_paths=20000, n_steps=100, seed=42).price(call, st, hes, rt).price
    g["heston_v0kw_0.09_4000x20_seed1"] = MonteCarloTechnique(n_paths=4000, n_steps=20, seed=1).price(call, st, hes, rt, v0=0.09).price
    g["heston_fft"] = FFTTechnique().price(call, st, hes, rt, v0=0.04).price
    sabr = SABRModel(params=SABRModel.default_params)
    g["G3_sabr_20000x100_seed42"] = MonteCarloTechnique(n_paths=20000, n_steps=100, seed=42).price(call, st, sabr, rt).price
    for name, mdl in (("merton", MertonJumpModel(params=MertonJumpModel.default_params)),
                      ("bates", BatesModel(params=BatesModel.default_params)),
                      ("kou", KouModel(params=KouModel.default_params)),
                      ("sabr_jump", SABRJumpModel(params=SABRJumpModel.default_params))):
        seed_numba(1234)
        g[f"G4_{name}_20000x100_seed42_numba1234"] = MonteCarloTechnique(n_paths=20000, n_steps=100, seed=42).price(call, st, mdl, rt).price
    g["merton_closed_form"] = ClosedFormTechnique().price(call, st, MertonJumpModel(params=MertonJumpModel.default_params), rt).price
    g["bates_fft"] = FFTTechnique().price(call, st, BatesModel(params=BatesModel.default_params), rt, v0=0.04).price
    g["kou_fft"] = FFTTechnique().price(call, st, KouModel(params=KouModel.default_params), rt).price
    aput = Option(strike=110.0, maturity=1.0, option_type=OptionType.PUT)
    g["G5_lsmc_bsm_50000x100_seed42_deg2"] = float(AmericanMonteCarloTechnique(n_paths=50000, n_steps=100, seed=42).price(aput, stq, bsm, rt).price)
    g["G6_lsmc_bsm_50000x50_seed42_deg3"] = float(AmericanMonteCarloTechnique(n_paths=50000, n_steps=50, seed=42, lsm_degree=3).price(aput, stq, bsm, rt).price)
    g["lsmc_heston_8000x25_seed9_deg2"] = float(AmericanMonteCarloTechnique(n_paths=8000, n_steps=25, seed=9).price(aput, stq, hes, rt).price)
    seed_numba(99)
    g["lsmc_merton_8000x25_seed9_deg2_numba99"] = float(AmericanMonteCarloTechnique(n_paths=8000, n_steps=25, seed=9).price(aput, stq, MertonJumpModel(params=MertonJumpModel.default_params), rt).price)
    g["crr1001_american_put"] = CRRTechnique(steps=1001, is_american=True).price(aput, stq, bsm, rt).price
    t = MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0)
    g["G7_bsm_delta"] = t.delta(call, st, bsm, rt)
    g["G7_bsm_gamma"] = t.gamma(call, st, bsm, rt)
    g["G7_bsm_vega"] = t.vega(call, st, bsm, rt)
    g["heston_vega_is_nan"] = bool(np.isnan(MonteCarloTechnique(n_paths=200, n_steps=5, seed=0).vega(call, st, hes, rt)))
    g["_toolchain"] = dict(numpy=np.__version__, numba=numba.__version__)
    with open(os.path.join(HERE, "price_goldens.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    return g


# ---------------------------------------------------------------- single-step models
def levy_goldens():
    """Pure-Levy / exact-sampler branches of MonteCarloTechnique.price (monte_carlo.py:103-106):
    the reference's seeded prices (pin the oracle restatement bit for bit), large-sample
    prices with standard errors and sample quantiles of the terminal spot (pin the device
    samplers statistically), and the reference's FFT prices where a cf exists."""
    from optpricing.models import CEVModel, CGMYModel, HyperbolicModel, NIGModel, VarianceGammaModel

    call = Option(strike=100, maturity=1.0, option_type=OptionType.CALL)
    put = Option(strike=95, maturity=0.5, option_type=OptionType.PUT)
    stq = Stock(spot=100.0, dividend=0.01)
    rt = Rate(rate=0.05)
    models = {
        "vg": VarianceGammaModel(params=VarianceGammaModel.default_params),
        "vg_small_shape": VarianceGammaModel(params={"sigma": 0.25, "nu": 1.5, "theta": -0.1}),
        "nig": NIGModel(params=NIGModel.default_params),
        "cgmy": CGMYModel(params={"C": 0.5, "G": 6.0, "M": 9.0, "Y": 1.000001}),
        "hyperbolic": HyperbolicModel(params=HyperbolicModel.default_params),
        "cev": CEVModel(params=CEVModel.default_params),
        "cev_gamma1": CEVModel(params={"sigma": 0.2, "gamma": 1.0}),
    }
    qs = [0.001, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999]
    g = {"_quantile_levels": qs, "_toolchain": dict(numpy=np.__version__)}
    import scipy
    g["_toolchain"]["scipy"] = scipy.__version__
    for name, mdl in models.items():
        e = {"params": dict(mdl.params)}
        for tag, opt in (("call", call), ("put", put)):
            np.random.seed(4321)                       # CEV draws from numpy's global state
            e[f"{tag}_20000_seed42"] = MonteCarloTechnique(n_paths=20000, seed=42).price(opt, stq, mdl, rt).price
        # large sample: price, standard error, quantiles of S_T (call maturity)
        np.random.seed(99)
        t = MonteCarloTechnique(n_paths=2_000_000, seed=7)
        if name.startswith("cev"):
            ST = mdl.sample_terminal_spot(100.0, 0.05, 1.0, t.n_paths)
        else:
            ST = t._simulate_levy_terminal(mdl, 100.0, 0.05, 0.01, 1.0)
        pay = np.maximum(ST - 100.0, 0) * math.exp(-0.05)
        e["call_2m_price"] = float(pay.mean())
        e["call_2m_stderr"] = float(pay.std(ddof=1) / math.sqrt(pay.size))
        e["ST_2m_mean"] = float(ST.mean())
        e["ST_2m_quantiles"] = [float(x) for x in np.quantile(ST, qs)]
        if not name.startswith("cev"):
            try:
                e["call_fft"] = float(FFTTechnique().price(call, stq, mdl, rt).price)
            except Exception as exc:                  # noqa: BLE001
                e["call_fft_error"] = repr(exc)
        g[name] = e
    with open(os.path.join(HERE, "levy_goldens.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    return g


# ---------------------------------------------------------------- European, all seven models, ragged
def european_goldens():
    """MonteCarloTechnique.price for every SDE model on awkward shapes: a put with a dividend
    yield, odd n_paths with antithetic (rounded up, monte_carlo.py:58-59) and without."""
    put = Option(strike=105.0, maturity=0.5, option_type=OptionType.PUT)
    stq = Stock(spot=100.0, dividend=0.02)
    rt = Rate(rate=0.03)
    models = {
        "bsm": BSMModel(params={"sigma": 0.3}),
        "heston": HestonModel(params=HestonModel.default_params),
        "merton": MertonJumpModel(params=MertonJumpModel.default_params),
        "bates": BatesModel(params=BatesModel.default_params),
        "kou": KouModel(params=KouModel.default_params),
        "sabr": SABRModel(params=SABRModel.default_params),
        "sabr_jump": SABRJumpModel(params=SABRJumpModel.default_params),
    }
    g = {"_toolchain": dict(numpy=np.__version__, numba=numba.__version__)}
    for name, mdl in models.items():
        for tag, kw in (("anti_7001x13", dict(n_paths=7001, n_steps=13, antithetic=True)),
                        ("noanti_5001x13", dict(n_paths=5001, n_steps=13, antithetic=False))):
            seed_numba(55)
            g[f"{name}_put_{tag}_seed21"] = float(
                MonteCarloTechnique(seed=21, **kw).price(put, stq, mdl, rt).price)
    with open(os.path.join(HERE, "european_goldens.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    return g


# ---------------------------------------------------------------- greeks (GreekMixin on CRN)
def greek_goldens():
    """delta / gamma / vega / theta / rho of greek_mixin.py:29-265 for a European BSM put and a
    Heston call (models without jumps: the re-pricings are reproducible from the Generator
    alone), and delta / rho of an American put."""
    put = Option(strike=105.0, maturity=0.5, option_type=OptionType.PUT)
    call = Option(strike=100.0, maturity=1.0, option_type=OptionType.CALL)
    stq = Stock(spot=100.0, dividend=0.02)
    rt = Rate(rate=0.03)
    g = {"_toolchain": dict(numpy=np.__version__, numba=numba.__version__)}
    for name, opt, mdl in (("bsm_put", put, BSMModel(params={"sigma": 0.3})),
                           ("heston_call", call, HestonModel(params=HestonModel.default_params))):
        t = MonteCarloTechnique(n_paths=8000, n_steps=12, seed=33)
        for greek in ("delta", "gamma", "vega", "theta", "rho"):
            v = getattr(t, greek)(opt, stq, mdl, rt)
            g[f"{name}_{greek}"] = None if np.isnan(v) else float(v)
    a = AmericanMonteCarloTechnique(n_paths=6000, n_steps=15, seed=9)
    aput = Option(strike=105.0, maturity=0.75, option_type=OptionType.PUT)
    for greek in ("delta", "rho"):
        g[f"american_bsm_put_{greek}"] = float(getattr(a, greek)(aput, stq, BSMModel(params={"sigma": 0.25}), rt))
    with open(os.path.join(HERE, "greek_goldens.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    return g


# ---------------------------------------------------------------- American, all seven models
def american_goldens():
    """AmericanMonteCarloTechnique.price for every SDE model (american_monte_carlo.py:52-207),
    seeded: Generator seed 9 for the increments / counts, numba's global RNG seeded per model
    for the jump sizes.  Pins rng_mode="numpy" + the fed path kernels + the LS pricer together."""
    aput = Option(strike=105.0, maturity=0.75, option_type=OptionType.PUT)
    acall = Option(strike=95.0, maturity=0.75, option_type=OptionType.CALL)
    stq = Stock(spot=100.0, dividend=0.03)
    rt = Rate(rate=0.04)
    models = {
        "bsm": BSMModel(params={"sigma": 0.25}),
        "heston": HestonModel(params=HestonModel.default_params),
        "merton": MertonJumpModel(params=MertonJumpModel.default_params),
        "bates": BatesModel(params=BatesModel.default_params),
        "kou": KouModel(params=KouModel.default_params),
        "sabr": SABRModel(params=SABRModel.default_params),
        "sabr_jump": SABRJumpModel(params=SABRJumpModel.default_params),
    }
    g = {"_toolchain": dict(numpy=np.__version__, numba=numba.__version__)}
    for name, mdl in models.items():
        for tag, opt, deg in (("put_deg2", aput, 2), ("call_deg3", acall, 3)):
            seed_numba(77)
            t = AmericanMonteCarloTechnique(n_paths=6000, n_steps=15, seed=9, lsm_degree=deg)
            g[f"{name}_{tag}"] = float(t.price(opt, stq, mdl, rt).price)
    with open(os.path.join(HERE, "american_goldens.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    return g


# ---------------------------------------------------------------- Dupire
def dupire_vectors():
    """dupire_kernel (mc_kernels.py:340-359) cannot be compiled by numba with a Python
    surface; its un-jitted body, dupire_kernel.py_func, is the reference's own code."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.reference_mc import dupire_surfaces
    rng = np.random.default_rng(2024)
    n, m, dt = 257, 37, 1.0 / 37
    dw = rng.standard_normal((n, m)) * math.sqrt(dt)
    out = {"dw": dw, "meta": json.dumps(dict(n=n, m=m, dt=dt, log_s0=math.log(100.0), r=0.05, q=0.01))}
    for name, f in dupire_surfaces().items():
        out[name] = mc_kernels.dupire_kernel.py_func(n, m, math.log(100.0), 0.05, 0.01, dt, dw, f)
    np.savez_compressed(os.path.join(HERE, "dupire_vectors.npz"), **out)
    return out


if __name__ == "__main__":
    if "--greeks-only" in sys.argv:
        for k, v in greek_goldens().items():
            print(k, v)
        raise SystemExit(0)
    if "--european-only" in sys.argv:
        for k, v in european_goldens().items():
            print(k, v)
        raise SystemExit(0)
    if "--american-only" in sys.argv:
        for k, v in american_goldens().items():
            print(k, v)
        raise SystemExit(0)
    if "--dupire-only" in sys.argv:
        print("dupire_vectors.npz:", sorted(dupire_vectors()))
        raise SystemExit(0)
    if "--levy-only" in sys.argv:
        for k, v in levy_goldens().items():
            print(k, v)
        raise SystemExit(0)
    kv = kernel_vectors()
    print("kernel_vectors.npz:", len(kv), "arrays")
    print("ls_vectors.npz:", len(ls_vectors()), "cases")
    for k, v in price_goldens().items():
        print(k, v)
    for k, v in levy_goldens().items():
        print(k, v)
    print("dupire_vectors.npz:", sorted(dupire_vectors()))
    for k, v in american_goldens().items():
        print(k, v)
    for k, v in european_goldens().items():
        print(k, v)
    for k, v in greek_goldens().items():
        print(k, v)
