"""
Parity of the CUDA path against the oracle and the reference's golden vectors.
Everything here goes through the C ABI of liboptmc.so (via optpricing_b200) on
a real B200.  The oracle (oracle/) is only the checker.

Bars (BASELINE.json north_star):
  * fed the reference's own draws: per-path values and prices within 1e-12
    relative of the numba reference (fixtures) / the oracle (larger seeded cases);
  * on the device RNG: price within 3 combined standard errors of the oracle's
    price on its own draws, and of the closed form where one exists.
"""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import optpricing_b200 as ob  # noqa: E402
from optpricing_b200 import _engine  # noqa: E402
from optpricing_b200.techniques.kernels import (american_mc_kernels, mc_kernels,  # noqa: E402
                                                path_kernels)
from oracle import reference_mc as orc  # noqa: E402

KINDS = ("bsm", "heston", "merton", "bates", "sabr", "sabr_jump", "kou")
RTOL = 1e-12
MODELS = {"bsm": ob.BSMModel, "heston": ob.HestonModel, "merton": ob.MertonJumpModel,
          "bates": ob.BatesModel, "kou": ob.KouModel, "sabr": ob.SABRModel,
          "sabr_jump": ob.SABRJumpModel}
TERMINAL = {k: getattr(mc_kernels, f"{k}_kernel") for k in KINDS}
PATHS = {k: getattr(path_kernels, f"{k}_path_kernel") for k in KINDS}


@pytest.fixture(scope="module")
def eng():
    return _engine.get_engine()


# ------------------------------------------------------------------ RNG
def test_philox_matches_known_answers(eng):
    from philox_ref import KAT, philox4x32_10
    n = 1000
    seed = 0x299F31D0A4093822
    got = eng.philox_raw(seed, n, c2=0x13198A2E, c3=0x03707344)
    ids = np.arange(n, dtype=np.uint64)
    want = philox4x32_10(ids.astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                         np.full(n, 0x13198A2E, np.uint32), np.full(n, 0x03707344, np.uint32),
                         seed & 0xFFFFFFFF, seed >> 32)
    assert np.array_equal(got, want)
    got0 = eng.philox_raw(0, 1)[0]
    assert [int(x) for x in got0] == list(KAT[0][2])


@pytest.mark.parametrize("impl", (0, 1))
def test_device_normals_are_standard_normal(eng, impl):
    from scipy import stats
    eng.set_int("normals_impl", impl)
    try:
        z = eng.normals(seed=12345, n=1 << 21, step=3)
    finally:
        eng.set_int("normals_impl", 1)
    n = z.size
    assert abs(z.mean()) < 5 / math.sqrt(n)
    assert abs(z.var() - 1) < 5 * math.sqrt(2 / n)
    assert abs(stats.skew(z)) < 5 * math.sqrt(6 / n)
    assert abs(stats.kurtosis(z)) < 5 * math.sqrt(24 / n)
    a, b = z[0::2], z[1::2]
    assert abs(np.corrcoef(a, b)[0, 1]) < 5 / math.sqrt(n / 2)
    assert stats.kstest(z[: 1 << 18], "norm").pvalue > 1e-3
    # tails: P(|z| > 4) = 6.33e-5
    frac = np.mean(np.abs(z) > 4)
    assert abs(frac - 6.334e-5) < 5 * math.sqrt(6.334e-5 / n)
    # the two draws of a pair lie on a circle of the radius the first 64 bits give
    assert np.isfinite(z).all()


def test_normals_depend_only_on_seed_draw_step(eng):
    a = eng.normals(seed=7, n=4096, step=5)
    b = eng.normals(seed=7, n=1024, step=5)
    assert np.array_equal(a[:2048], b)
    assert not np.array_equal(a, eng.normals(seed=8, n=4096, step=5))
    assert not np.array_equal(a, eng.normals(seed=7, n=4096, step=6))


# ------------------------------------------------------------------ fed kernels
def _golden_args(z, meta, case, kind):
    kw = dict(n_paths=meta["D"], n_steps=meta["M"], **meta["common"], **meta["models"][kind])
    for k in ("dw", "dw1", "dw2", "jump_counts"):
        if f"{case}/{kind}/{k}" in z:
            kw[k] = z[f"{case}/{kind}/{k}"]
    return kw


@pytest.mark.parametrize("case", ("rand", "zero"))
@pytest.mark.parametrize("kind", KINDS)
def test_fed_kernels_match_reference_vectors(kernel_vectors, case, kind):
    z, meta = kernel_vectors
    kw = _golden_args(z, meta, case, kind)
    seed = int(z[f"{case}/{kind}/jump_seed"]) if f"{case}/{kind}/jump_seed" in z else 0
    np.random.seed(seed)     # numpy's legacy stream == numba's (make_golden.py asserts it)
    got = TERMINAL[kind](**kw)
    np.testing.assert_allclose(got, z[f"{case}/{kind}/terminal"], rtol=RTOL, atol=0)
    np.random.seed(seed)
    got = PATHS[kind](**kw)
    assert got.shape == (meta["D"], meta["M"] + 1)
    np.testing.assert_allclose(got, z[f"{case}/{kind}/paths"], rtol=RTOL, atol=0)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shape", ((1, 1), (63, 31), (65, 33), (1000, 70), (200003, 37)))
def test_fed_kernels_match_oracle_ragged_shapes(kernel_vectors, kind, shape):
    """Tile edges: path counts and step counts around the 64 x 32 staging tile."""
    _, meta = kernel_vectors
    n, m = shape
    rng = np.random.default_rng(n * 1000 + m)
    sc = dict(meta["models"][kind])
    kw = dict(n_paths=n, n_steps=m, **meta["common"], **sc)
    sd = math.sqrt(meta["common"]["dt"])
    if kind in orc.VARIANCE_KINDS:
        kw["dw1"], kw["dw2"] = rng.standard_normal((n, m)) * sd, rng.standard_normal((n, m)) * sd
    else:
        kw["dw"] = rng.standard_normal((n, m)) * sd
    jb = None
    if kind in orc.JUMP_KINDS:
        kw["jump_counts"] = rng.poisson(0.4, (n, m))
        jb = orc.draw_jump_buffer(kind, kw, np.random.RandomState(99))
    want_t = orc.run_kernel(kind, kw, jump_buf=jb)
    want_p = orc.run_kernel(kind, kw, jump_buf=jb, want_paths=True)
    np.random.seed(99)
    np.testing.assert_allclose(TERMINAL[kind](**kw), want_t, rtol=RTOL, atol=0)
    np.random.seed(99)
    np.testing.assert_allclose(PATHS[kind](**kw), want_p, rtol=RTOL, atol=0)


# ------------------------------------------------------------------ technique, reference draws
CALL = ob.Option(strike=100, maturity=1.0, option_type=ob.OptionType.CALL)
PUT = ob.Option(strike=100, maturity=1.0, option_type=ob.OptionType.PUT)
ST, STQ, RT = ob.Stock(spot=100.0), ob.Stock(spot=100.0, dividend=0.01), ob.Rate(rate=0.05)


def test_numpy_mode_reproduces_reference_prices(price_goldens):
    g = price_goldens
    T = ob.MonteCarloTechnique
    bsm = ob.BSMModel({"sigma": 0.2})
    assert T(n_paths=20000, n_steps=10, seed=0, rng_mode="numpy").price(CALL, ST, bsm, RT).price \
        == pytest.approx(g["G1_bsm_20000x10_seed0"], rel=RTOL)
    assert T(n_paths=19999, n_steps=7, seed=3, rng_mode="numpy").price(PUT, STQ, bsm, RT).price \
        == pytest.approx(g["bsm_put_q_19999x7_seed3"], rel=RTOL)
    assert T(n_paths=5000, n_steps=9, antithetic=False, seed=5, rng_mode="numpy").price(
        CALL, ST, bsm, RT).price == pytest.approx(g["bsm_noanti_5000x9_seed5"], rel=RTOL)
    assert T(n_paths=20000, n_steps=100, seed=42, rng_mode="numpy").price(
        CALL, ST, ob.HestonModel(), RT).price == pytest.approx(g["G2_heston_20000x100_seed42"], rel=RTOL)
    assert T(n_paths=4000, n_steps=20, seed=1, rng_mode="numpy").price(
        CALL, ST, ob.HestonModel(), RT, v0=0.09).price == pytest.approx(
            g["heston_v0kw_0.09_4000x20_seed1"], rel=RTOL)
    assert T(n_paths=20000, n_steps=100, seed=42, rng_mode="numpy").price(
        CALL, ST, ob.SABRModel(), RT).price == pytest.approx(g["G3_sabr_20000x100_seed42"], rel=RTOL)
    for kind in ("merton", "bates", "kou", "sabr_jump"):
        np.random.seed(1234)
        got = T(n_paths=20000, n_steps=100, seed=42, rng_mode="numpy").price(
            CALL, ST, MODELS[kind](), RT).price
        assert got == pytest.approx(g[f"G4_{kind}_20000x100_seed42_numba1234"], rel=RTOL), kind


def test_numpy_mode_reproduces_reference_greeks(price_goldens):
    g = price_goldens
    t = ob.MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0, rng_mode="numpy")
    bsm = ob.BSMModel({"sigma": 0.2})
    assert t.delta(CALL, ST, bsm, RT) == pytest.approx(g["G7_bsm_delta"], rel=1e-9)
    assert t.gamma(CALL, ST, bsm, RT) == pytest.approx(g["G7_bsm_gamma"], rel=1e-6)
    assert t.vega(CALL, ST, bsm, RT) == pytest.approx(g["G7_bsm_vega"], rel=1e-9)
    assert np.isnan(t.vega(CALL, ST, ob.HestonModel(), RT))


def test_numpy_mode_reproduces_all_five_reference_greeks(greek_goldens):
    """delta / gamma / vega / theta / rho (greek_mixin.py:29-265) of a European BSM put and a
    Heston call, and delta / rho of an American put, from the reference's seeds."""
    g = greek_goldens
    put = ob.Option(strike=105.0, maturity=0.5, option_type=ob.OptionType.PUT)
    stock, rate = ob.Stock(spot=100.0, dividend=0.02), ob.Rate(rate=0.03)
    tol = {"delta": 1e-8, "gamma": 1e-4, "vega": 1e-8, "theta": 1e-6, "rho": 1e-8}
    for name, opt, model in (("bsm_put", put, ob.BSMModel({"sigma": 0.3})),
                             ("heston_call", CALL, ob.HestonModel())):
        t = ob.MonteCarloTechnique(n_paths=8000, n_steps=12, seed=33, rng_mode="numpy")
        for greek, rel in tol.items():
            got, want = getattr(t, greek)(opt, stock, model, rate), g[f"{name}_{greek}"]
            if want is None:
                assert math.isnan(got), (name, greek)
            else:
                assert got == pytest.approx(want, rel=rel), (name, greek)
    a = ob.AmericanMonteCarloTechnique(n_paths=6000, n_steps=15, seed=9, rng_mode="numpy")
    aput = ob.Option(strike=105.0, maturity=0.75, option_type=ob.OptionType.PUT)
    for greek in ("delta", "rho"):
        got = getattr(a, greek)(aput, stock, ob.BSMModel({"sigma": 0.25}), rate)
        assert got == pytest.approx(g[f"american_bsm_put_{greek}"], rel=1e-6), greek


# ------------------------------------------------------------------ LSMC
@pytest.fixture
def lsmc_mode(request, eng):
    """1 = all dates in one persistent launch (default), 0 = one launch per date."""
    eng.set_int("lsmc_mode", request.param)
    yield request.param
    eng.set_int("lsmc_mode", 1)


@pytest.mark.parametrize("lsmc_mode", [1, 0], indirect=True)
def test_ls_pricer_matches_reference_vectors(ls_vectors, lsmc_mode):
    z, cases = ls_vectors
    paths, dt, r = z["paths"], float(z["dt"]), float(z["r"])
    for c in cases:
        p = np.ascontiguousarray(paths[:, :2]) if c.get("one_step") else paths
        got = american_mc_kernels.longstaff_schwartz_pricer(p, c["K"], dt, r, c["is_call"],
                                                            c["degree"])
        if c["K"] == 60.0 and c["degree"] == 3:
            # rank-deficient dates: the reference's value is LAPACK rounding noise
            # (see tests/test_oracle_golden.py); the engine skips such dates.
            assert got == pytest.approx(c["price"], rel=0.5)
            continue
        assert got == pytest.approx(c["price"], rel=1e-10, abs=1e-13), c


def test_american_numpy_mode_reproduces_reference_prices(price_goldens):
    g = price_goldens
    A = ob.AmericanMonteCarloTechnique
    aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
    bsm = ob.BSMModel({"sigma": 0.2})
    got = A(n_paths=50000, n_steps=100, seed=42, rng_mode="numpy").price(aput, STQ, bsm, RT).price
    assert got == pytest.approx(g["G5_lsmc_bsm_50000x100_seed42_deg2"], rel=1e-10)
    assert got == pytest.approx(g["crr1001_american_put"], abs=0.1)   # the reference's own test
    got = A(n_paths=50000, n_steps=50, seed=42, lsm_degree=3, rng_mode="numpy").price(
        aput, STQ, bsm, RT).price
    assert got == pytest.approx(g["G6_lsmc_bsm_50000x50_seed42_deg3"], rel=1e-10)
    got = A(n_paths=8000, n_steps=25, seed=9, rng_mode="numpy").price(
        aput, STQ, ob.HestonModel(), RT).price
    assert got == pytest.approx(g["lsmc_heston_8000x25_seed9_deg2"], rel=1e-10)
    np.random.seed(99)
    got = A(n_paths=8000, n_steps=25, seed=9, rng_mode="numpy").price(
        aput, STQ, ob.MertonJumpModel(), RT).price
    assert got == pytest.approx(g["lsmc_merton_8000x25_seed9_deg2_numba99"], rel=1e-10)


@pytest.mark.parametrize("n_paths,n_steps,degree", [(1, 3, 2), (7, 1, 1), (1001, 5, 3), (40000, 12, 3),
                                                    (300001, 7, 4), (20000, 6, 7)])
def test_lsmc_persistent_launch_equals_per_date_launches(eng, n_paths, n_steps, degree):
    """Same arithmetic per path, different summation order of the power sums:
    the two drivers agree to rounding, and both with the oracle on the same paths."""
    rng = np.random.default_rng(n_paths + degree)
    z = rng.standard_normal((n_paths, n_steps)) * 0.2 * math.sqrt(1.0 / n_steps)
    paths = 100.0 * np.exp(np.concatenate([np.zeros((n_paths, 1)), np.cumsum(z, axis=1)], axis=1))
    got = {}
    for mode in (1, 0):
        eng.set_int("lsmc_mode", mode)
        try:
            got[mode] = american_mc_kernels.longstaff_schwartz_pricer(paths, 105.0, 1.0 / n_steps,
                                                                      0.05, False, degree)
        finally:
            eng.set_int("lsmc_mode", 1)
    assert got[1] == pytest.approx(got[0], rel=1e-9 if degree >= 4 else 1e-11, abs=1e-13)
    # the sweep fed by the cp.async ring (taken from 2^19 paths up by default): forced on here --
    # same paths per thread, same order of the sums, so the very same number
    eng.set_int("lsmc_ring", 1)
    try:
        ring = american_mc_kernels.longstaff_schwartz_pricer(paths, 105.0, 1.0 / n_steps, 0.05,
                                                             False, degree)
    finally:
        eng.set_int("lsmc_ring", 2)
    assert ring == got[1]
    if degree <= 3 and n_paths > 100:
        want = orc.longstaff_schwartz(paths, 105.0, 1.0 / n_steps, 0.05, False, degree)
        assert got[1] == pytest.approx(want, rel=1e-9)


@pytest.mark.parametrize("kind", KINDS)
def test_european_numpy_mode_ragged_all_models(european_goldens, kind):
    """rng_mode="numpy" on awkward shapes (put, dividend, odd n_paths, with / without antithetic)
    for every SDE model: the reference's price to 1e-12."""
    model = MODELS[kind]() if kind != "bsm" else ob.BSMModel({"sigma": 0.3})
    put = ob.Option(strike=105.0, maturity=0.5, option_type=ob.OptionType.PUT)
    stock, rate = ob.Stock(spot=100.0, dividend=0.02), ob.Rate(rate=0.03)
    for tag, n, anti in (("anti_7001x13", 7001, True), ("noanti_5001x13", 5001, False)):
        np.random.seed(55)
        t = ob.MonteCarloTechnique(n_paths=n, n_steps=13, antithetic=anti, seed=21, rng_mode="numpy")
        got = t.price(put, stock, model, rate).price
        assert got == pytest.approx(european_goldens[f"{kind}_put_{tag}_seed21"], rel=RTOL), (kind, tag)


@pytest.mark.parametrize("kind", KINDS)
def test_american_numpy_mode_all_models(american_goldens, kind):
    """rng_mode="numpy": fed path kernels + persistent LS pricer reproduce the reference's
    American price for every SDE model (put degree 2, call degree 3) from the same seeds."""
    model = MODELS[kind]() if kind != "bsm" else ob.BSMModel({"sigma": 0.25})
    stock, rate = ob.Stock(spot=100.0, dividend=0.03), ob.Rate(rate=0.04)
    for tag, K, typ, deg in (("put_deg2", 105.0, ob.OptionType.PUT, 2),
                             ("call_deg3", 95.0, ob.OptionType.CALL, 3)):
        np.random.seed(77)
        t = ob.AmericanMonteCarloTechnique(n_paths=6000, n_steps=15, seed=9, lsm_degree=deg,
                                           rng_mode="numpy")
        got = t.price(ob.Option(strike=K, maturity=0.75, option_type=typ), stock, model, rate).price
        assert got == pytest.approx(american_goldens[f"{kind}_{tag}"], rel=1e-8), (kind, tag)


def test_lsmc_high_degree_is_statistically_consistent():
    """degree 5: the reference itself is off by ~3e-3 from the stable solve."""
    rng = np.random.default_rng(11)
    paths = orc.simulate_paths("bsm", {"sigma": 0.2}, 100.0, 0.05, 0.01, 1.0, 20000, 20, True, rng)
    want = orc.longstaff_schwartz(paths, 110.0, 0.05, 0.05, False, 5)
    got = american_mc_kernels.longstaff_schwartz_pricer(paths, 110.0, 0.05, 0.05, False, 5)
    assert got == pytest.approx(want, abs=0.02)


# ------------------------------------------------------------------ device RNG: statistics
def _oracle_price_and_se(kind, p, S0, K, T, r, q, call, n_paths, n_steps, seed):
    price, STv = orc.price_european(kind, p, S0, K, T, r, q, call, n_paths, n_steps, True,
                                    np.random.default_rng(seed),
                                    jump_rs=np.random.RandomState(seed), return_terminal=True)
    pay = np.maximum(STv - K, 0) if call else np.maximum(K - STv, 0)
    half = pay.size // 2
    pair = 0.5 * (pay[:half] + pay[half:])
    return price, math.exp(-r * T) * pair.std(ddof=1) / math.sqrt(half)


@pytest.mark.parametrize("impl", (0, 1))
@pytest.mark.parametrize("kind", KINDS)
def test_philox_price_within_3_sigma_of_oracle(eng, kind, impl):
    S0, K, T, r, q, n_steps = 100.0, 100.0, 1.0, 0.05, 0.01, 50
    mdl = MODELS[kind]()
    want, se_ref = _oracle_price_and_se(kind, mdl.params, S0, K, T, r, q, True, 200_000, n_steps, 5)
    eng.set_int("normals_impl", impl)
    try:
        t = ob.MonteCarloTechnique(n_paths=1 << 21, n_steps=n_steps, seed=17)
        got = t.price(ob.Option(K, T, ob.OptionType.CALL), ob.Stock(S0, dividend=q), mdl,
                      ob.Rate(r)).price
    finally:
        eng.set_int("normals_impl", 1)
    se = math.hypot(t.last_stats["stderr"], se_ref)
    assert abs(got - want) < 3 * se, (kind, got, want, se)


def test_philox_bsm_against_closed_form(price_goldens):
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=10, seed=0)
    got = t.price(CALL, ST, ob.BSMModel({"sigma": 0.2}), RT).price
    assert abs(got - price_goldens["G1_closed_form"]) < 3.5 * t.last_stats["stderr"]
    # The reference's own acceptance test (tests/techniques/test_monte_carlo.py:111-126) asserts
    # abs=1e-2 at 20 000 x 10, seed 0, where one standard error is 0.07: it holds for the draws
    # seed 0 happens to give numpy's generator.  rng_mode="numpy" reproduces those draws (G1 to
    # 1e-12), so it passes there by construction; on the device RNG the same contract is met
    # statistically -- within 3.5 standard errors -- and no tighter.
    t = ob.MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0, rng_mode="numpy")
    assert t.price(CALL, ST, ob.BSMModel({"sigma": 0.2}), RT).price == pytest.approx(
        price_goldens["G1_closed_form"], abs=1e-2)
    t = ob.MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0)
    got = t.price(CALL, ST, ob.BSMModel({"sigma": 0.2}), RT).price
    assert abs(got - price_goldens["G1_closed_form"]) < 3.5 * t.last_stats["stderr"]


def test_philox_independent_correlation_matches_fft(price_goldens):
    """Properly correlated drivers: Heston MC vs the reference's FFT price, allowing
    for the Euler bias of the absorbing scheme (+0.03 at 252 steps)."""
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=252, seed=3, correlation="independent")
    got = t.price(CALL, ST, ob.HestonModel(), RT).price
    assert abs(got - price_goldens["heston_fft"]) < 0.06 + 3 * t.last_stats["stderr"]
    # and the reference-compatible (double-correlated) mode is measurably different
    t2 = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=252, seed=3)
    ref_mode = t2.price(CALL, ST, ob.HestonModel(), RT).price
    assert got - ref_mode > 0.08


@pytest.mark.parametrize("kind", KINDS)
def test_control_mean_is_the_scheme_martingale(kind):
    """E[S_T] of the discrete scheme in closed form, at BASELINE size for Heston."""
    n_paths = 1 << 24 if kind == "heston" else 1 << 21
    n_steps = 252 if kind == "heston" else 40
    t = ob.MonteCarloTechnique(n_paths=n_paths, n_steps=n_steps, seed=23)
    t.price(CALL, STQ, MODELS[kind](), RT)
    s = t.last_stats
    n = s["n_samples"]
    assert n == n_paths // 2
    assert 0 < s["control_stderr"] < 40.0 / math.sqrt(n)
    assert abs(s["control_mean"] - s["control_expected"]) < 4 * s["control_stderr"]


def test_put_call_parity_holds_per_sample():
    """Same seed => same paths: C - P = disc * (mean S_T - K) to rounding, 2^24 x 252."""
    t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=252, seed=5)
    c = t.price(CALL, ST, ob.HestonModel(), RT).price
    mean_c = t.last_stats["control_mean"]
    t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=252, seed=5)
    p = t.price(PUT, ST, ob.HestonModel(), RT).price
    assert c - p == pytest.approx(math.exp(-0.05) * (mean_c - 100.0), rel=1e-10)


def test_seed_semantics_match_reference():
    """seed= reproduces; successive price() calls advance; crn() replays."""
    mk = lambda: ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=8, seed=9)  # noqa: E731
    a, b = mk(), mk()
    m = ob.BSMModel()
    p1, p2 = a.price(CALL, ST, m, RT).price, a.price(CALL, ST, m, RT).price
    assert p1 != p2
    assert b.price(CALL, ST, m, RT).price == p1
    with ob.crn(b.rng):
        q1 = b.price(CALL, ST, m, RT).price
    assert b.price(CALL, ST, m, RT).price == q1 == p2


def test_philox_greeks_on_common_random_numbers(price_goldens):
    t = ob.MonteCarloTechnique(n_paths=1 << 20, n_steps=10, seed=0)
    bsm = ob.BSMModel({"sigma": 0.2})
    assert t.delta(CALL, ST, bsm, RT) == pytest.approx(0.63683, abs=2e-3)
    assert t.gamma(CALL, ST, bsm, RT) == pytest.approx(0.018762, abs=3e-3)
    assert t.vega(CALL, ST, bsm, RT) == pytest.approx(37.524, abs=0.3)
    assert np.isnan(t.vega(CALL, ST, ob.HestonModel(), RT))
    # jump models too: jump sizes are Philox draws, so CRN holds (the reference's are not)
    d = ob.MonteCarloTechnique(n_paths=1 << 20, n_steps=20, seed=1).delta(
        CALL, ST, ob.MertonJumpModel(), RT)
    assert 0.55 < d < 0.75


def test_control_variate_reduces_error():
    t = ob.MonteCarloTechnique(n_paths=1 << 20, n_steps=20, seed=4, control_variate=True)
    deep = ob.Option(strike=60.0, maturity=1.0, option_type=ob.OptionType.CALL)
    t.price(deep, ST, ob.BSMModel(), RT)
    s = t.last_stats
    assert s["cv_stderr"] < 0.5 * s["stderr"]
    assert abs(s["cv_price"] - s["price"]) < 4 * s["stderr"]


def test_rank_partition_sums_to_the_whole(eng):
    """R ranks emulated sequentially on one GPU: moments add up to the 1-rank job
    (streams are keyed by global draw id)."""
    mdl = _engine.make_model("bates", ob.BatesModel().params, 0.05, 0.01)
    n = 100_001
    whole = eng.european(mdl, eng.make_sim(100.0, 1.0, n, 16, True, 77, 0), [100.0], [1])[0]
    for world in (2, 3, 8):
        acc = np.zeros(6)
        for rank in range(world):
            off, cnt = ob.techniques.monte_carlo.shard_draws(n, rank, world)
            acc += eng.european(mdl, eng.make_sim(100.0, 1.0, n, 16, True, 77, 0, off, cnt),
                                [100.0], [1])[0]
        np.testing.assert_allclose(acc, whole, rtol=1e-12)


def test_strike_sweep_matches_single_contract_calls(eng):
    mdl = _engine.make_model("bates", ob.BatesModel().params, 0.05, 0.01)
    sim = eng.make_sim(100.0, 0.5, 1 << 18, 30, True, 5, 0)
    ks = np.linspace(70, 130, 64)
    calls = np.arange(64) % 2
    sweep = eng.european(mdl, sim, ks, calls)
    assert sweep.shape == (64, 6)
    for j in (0, 17, 63):
        one = eng.european(mdl, sim, [ks[j]], [int(calls[j])])[0]
        np.testing.assert_allclose(sweep[j], one, rtol=1e-12)


def test_terminal_array_matches_reference_layout():
    t = ob.MonteCarloTechnique(n_paths=1000, n_steps=5, seed=2)
    STv = t._simulate_sde_path(ob.BSMModel(), 100.0, 0.05, 0.0, 1.0)
    assert isinstance(STv, np.ndarray) and STv.shape == (1000,)
    # antithetic partners: log-returns around the drift are mirrored
    drift = (0.05 - 0.5 * 0.04) * 1.0
    a, b = np.log(STv[:500] / 100.0) - drift, np.log(STv[500:] / 100.0) - drift
    np.testing.assert_allclose(a, -b, atol=1e-12)


def test_american_philox_price(price_goldens):
    """The reference's acceptance test (test_american_monte_carlo.py:20-39) on the device RNG,
    then BASELINE config 4's shape at a size the oracle can check statistically."""
    aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
    bsm = ob.BSMModel({"sigma": 0.2})
    got = ob.AmericanMonteCarloTechnique(n_paths=50000, n_steps=100, seed=42).price(
        aput, STQ, bsm, RT).price
    assert got == pytest.approx(price_goldens["crr1001_american_put"], abs=0.1)
    got = ob.AmericanMonteCarloTechnique(n_paths=1 << 20, n_steps=50, seed=1, lsm_degree=3).price(
        aput, STQ, bsm, RT).price
    assert got == pytest.approx(price_goldens["crr1001_american_put"], abs=0.05)
    m = np.asarray(ob.AmericanMonteCarloTechnique(n_paths=64, n_steps=6, seed=3)
                   ._simulate_full_sde_paths(aput, STQ, bsm, RT))
    assert m.shape == (64, 7) and np.all(m[:, 0] == 100.0) and np.all(m > 0)


@pytest.mark.parametrize("kind", KINDS)
def test_american_philox_all_models_run(kind):
    aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
    eu = ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=25, seed=8,
                                correlation="independent").price(aput, STQ, MODELS[kind](), RT).price
    am = ob.AmericanMonteCarloTechnique(n_paths=1 << 18, n_steps=25, seed=8).price(
        aput, STQ, MODELS[kind](), RT).price
    assert am > eu - 0.05 and am < eu + 5.0     # early-exercise premium is non-negative


@pytest.fixture(params=[1, 0], ids=["jump_schedule", "poisson_per_step"])
def jump_schedule(request, eng):
    """1 = the path kernels draw Poisson(lambda T) and the jump steps once (default),
    0 = one Poisson(lambda dt) draw per step."""
    eng.set_int("jump_schedule", request.param)
    yield request.param
    eng.set_int("jump_schedule", 1)


@pytest.mark.parametrize("kind", ("merton", "bates", "kou", "sabr_jump", "heston"))
def test_philox_path_store_law(kind, jump_schedule):
    """The device-RNG path kernels, date by date: the scheme's martingale mean at EVERY exercise
    date (jump timing and compensation), the antithetic layout, and the terminal column's
    put price against the oracle's path kernel within 3 combined standard errors."""
    from optpricing_b200.techniques.monte_carlo import control_mean
    model = MODELS[kind]()
    n, m, T, r, q = 1 << 19, 16, 1.0, 0.05, 0.01
    t = ob.AmericanMonteCarloTechnique(n_paths=n, n_steps=m, seed=14)
    opt = ob.Option(strike=105.0, maturity=T, option_type=ob.OptionType.PUT)
    paths = np.asarray(t._simulate_full_sde_paths(opt, STQ, model, RT))
    assert paths.shape == (n, m + 1) and np.all(paths[:, 0] == 100.0)
    pair = 0.5 * (paths[: n // 2] + paths[n // 2:])
    for i in (1, 2, m // 2, m):
        want = control_mean(kind, model.params, 100.0, r, q, T * i / m, i)
        se = pair[:, i].std(ddof=1) / math.sqrt(n // 2)
        assert abs(pair[:, i].mean() - want) < 4.5 * se, (kind, i, pair[:, i].mean(), want, se)
    pay = np.maximum(105.0 - paths[:, m], 0)
    got, se_g = pay.mean(), 0.5 * (pay[: n // 2] + pay[n // 2:]).std(ddof=1) / math.sqrt(n // 2)
    ref = orc.simulate_paths(kind, model.params, 100.0, r, q, T, 200_000, m, True,
                             np.random.default_rng(3), np.random.RandomState(4))
    pr = np.maximum(105.0 - ref[:, m], 0)
    se_r = 0.5 * (pr[:100_000] + pr[100_000:]).std(ddof=1) / math.sqrt(100_000)
    assert abs(got - pr.mean()) < 3 * math.hypot(se_g, se_r), (kind, got, pr.mean())
    # second moment of the log-return at the first date: the jump sizes' law
    lr = np.log(paths[:, 1] / 100.0)
    lr_ref = np.log(ref[:, 1] / 100.0)
    assert lr.var() == pytest.approx(lr_ref.var(), rel=0.05)


@pytest.mark.parametrize("kind", ("merton", "kou", "bates", "sabr_jump"))
def test_path_kernels_many_jumps_per_step_schedule_against_per_step_poisson(eng, kind):
    """Path kernels at lambda T = 2 on THREE dates: most paths jump, many twice on the same date
    (the pending jump is refilled and applied again within the step), and a few hundred of the
    2^20 paths have more than the eight jumps the schedule keeps in registers (the fallback that
    recomputes the indices).  The schedule and the per-step Poisson sampler are two samplers of
    one law: every date's mean, the second moment of every date's increment and the terminal put
    agree within 4.5 combined standard errors."""
    params = dict(MODELS[kind]().params)
    params["lambda"] = 2.0
    model = MODELS[kind](params=params)
    n, m = 1 << 20, 3
    opt = ob.Option(strike=105.0, maturity=1.0, option_type=ob.OptionType.PUT)
    out = {}
    try:
        for mode in (1, 0):
            eng.set_int("jump_schedule", mode)
            t = ob.AmericanMonteCarloTechnique(n_paths=n, n_steps=m, seed=23)
            out[mode] = np.asarray(t._simulate_full_sde_paths(opt, STQ, model, RT))
    finally:
        eng.set_int("jump_schedule", 1)
    half = n // 2

    def pair_stat(x):                       # mean and standard error over antithetic pairs
        pm = 0.5 * (x[:half] + x[half:])
        return pm.mean(), pm.std(ddof=1) / math.sqrt(half)

    for i in (1, 2, 3):
        # (squared increments, not log-returns: a SABR path may sit at its absorbing zero)
        for f in (lambda p: p[:, i], lambda p: ((p[:, i] - p[:, i - 1]) / 100.0) ** 2):
            (a, ea), (b, eb) = pair_stat(f(out[1])), pair_stat(f(out[0]))
            assert abs(a - b) < 4.5 * math.hypot(ea, eb), (kind, i, a, b, ea, eb)
    (a, ea), (b, eb) = (pair_stat(np.maximum(105.0 - out[k][:, m], 0.0)) for k in (1, 0))
    assert abs(a - b) < 4.5 * math.hypot(ea, eb), (kind, a, b)


@pytest.mark.parametrize("n_steps,antithetic,lam", [(37, True, 1.5), (64, False, 0.4), (3, True, 0.9),
                                                    (50, True, 0.05)])
def test_european_sabr_jump_schedule_matches_per_step_poisson(eng, n_steps, antithetic, lam):
    """European SABR-jump (jumps act on S inside the step loop): the jump schedule (next jump kept
    ready, applied branch-free) and the per-step Poisson draw are two samplers of one law -- prices, second moments and the
    martingale mean of S_T agree within 4 combined standard errors; ragged last groups, no
    antithetic partner, many and few jumps."""
    params = dict(ob.SABRJumpModel().params)
    params["lambda"] = lam
    model = ob.SABRJumpModel(params=params)
    out = {}
    try:
        for mode in (1, 0):
            eng.set_int("jump_schedule", mode)
            t = ob.MonteCarloTechnique(n_paths=1 << 21, n_steps=n_steps, antithetic=antithetic, seed=5)
            p = t.price(CALL, STQ, model, RT).price
            st = t.last_stats
            out[mode] = (p, st["stderr"], st["control_mean"], st["control_stderr"])
            fwd = st["control_expected"]          # the scheme's own E[S_T]
    finally:
        eng.set_int("jump_schedule", 1)
    (p1, e1, c1, s1), (p0, e0, c0, s0) = out[1], out[0]
    assert abs(p1 - p0) < 4 * math.hypot(e1, e0), out
    # the sample mean of S_T: the two samplers against each other and against the scheme's
    # own mean
    assert abs(c1 - c0) < 4 * math.hypot(s1, s0), out
    assert abs(c1 - fwd) < 4 * s1 + 1e-4 * fwd and abs(c0 - fwd) < 4 * s0 + 1e-4 * fwd, (out, fwd)


# ------------------------------------------------------------------ batched contracts, fused greeks
def test_price_many_equals_single_prices():
    """One simulation per maturity; each price equals price() on the same Philox key."""
    model = ob.BatesModel()
    strikes = np.linspace(70, 130, 64)
    opts = [ob.Option(float(k), T, ob.OptionType.CALL if i % 3 else ob.OptionType.PUT)
            for T in (0.25, 1.0) for i, k in enumerate(strikes)]
    t = ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=40, seed=21)
    many = t.price(opts, STQ, model, RT)
    assert isinstance(many, np.ndarray) and many.shape == (128,)
    seeds = {st["seed"] for st in t.last_stats_many}
    assert len(seeds) == 2                                   # one simulation per maturity
    # replay: a fresh technique draws the same first key; second maturity = second key
    t2 = ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=40, seed=21)
    first = t2.price(opts[5], STQ, model, RT).price
    assert first == pytest.approx(many[5], rel=1e-12)
    second = t2.price(opts[64 + 9], STQ, model, RT).price
    assert second == pytest.approx(many[64 + 9], rel=1e-12)
    # monotone in strike for calls of one maturity (same paths => exactly monotone)
    calls = [many[i] for i in range(64) if i % 3]
    assert all(a >= b for a, b in zip(calls, calls[1:]))
    # more than 64 contracts of one maturity: chunks reuse the stored terminals
    wide = [ob.Option(float(k), 1.0, ob.OptionType.CALL) for k in np.linspace(60, 140, 150)]
    t3 = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=10, seed=2)
    w = t3.price_many(wide, STQ, ob.BSMModel(), RT)
    t4 = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=10, seed=2)
    assert t4.price(wide[120], STQ, ob.BSMModel(), RT).price == pytest.approx(w[120], rel=1e-12)


def test_price_frame_is_the_mc_counterpart_of_the_vectorised_pricer():
    """DataFrame in (strike, maturity, optionType), prices out in row order -- the shape of
    price_options_vectorized (calibration/vectorized_pricer.py:20-28) -- for a model without
    a characteristic function."""
    import pandas as pd
    rng = np.random.default_rng(3)
    df = pd.DataFrame({"strike": rng.uniform(80, 120, 30), "maturity": rng.choice([0.5, 1.0, 2.0], 30),
                       "optionType": rng.choice(["call", "put"], 30)})
    df.index = rng.permutation(30) + 100                    # an arbitrary index must not matter
    model = ob.SABRModel()
    got = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=20, seed=4).price_frame(df, STQ, model, RT)
    opts = [ob.Option(float(k), float(t), ob.OptionType.CALL if c == "call" else ob.OptionType.PUT)
            for k, t, c in zip(df["strike"], df["maturity"], df["optionType"])]
    want = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=20, seed=4).price_many(opts, STQ, model, RT)
    assert got.shape == (30,) and np.array_equal(got, want)
    # BSM: the surface of implied vols recovered from the frame's prices is flat at sigma
    flat = pd.DataFrame({"strike": [90.0, 100.0, 110.0] * 2, "maturity": [0.5] * 3 + [1.0] * 3,
                         "optionType": ["call", "put", "call"] * 2})
    px = ob.MonteCarloTechnique(n_paths=1 << 20, n_steps=8, seed=9, control_variate=True).price_frame(
        flat, STQ, ob.BSMModel(params={"sigma": 0.2}), RT)
    from scipy.stats import norm

    def black_scholes(k, t, call, s0=100.0, r=0.05, q=0.01, sig=0.2):
        d1 = (math.log(s0 / k) + (r - q + 0.5 * sig * sig) * t) / (sig * math.sqrt(t))
        d2 = d1 - sig * math.sqrt(t)
        c = s0 * math.exp(-q * t) * norm.cdf(d1) - k * math.exp(-r * t) * norm.cdf(d2)
        return c if call else c - s0 * math.exp(-q * t) + k * math.exp(-r * t)

    for p, k, t, c in zip(px, flat["strike"], flat["maturity"], flat["optionType"]):
        assert p == pytest.approx(black_scholes(k, t, c == "call"), abs=0.02)


@pytest.mark.parametrize("kind", ("bsm", "heston", "merton", "bates", "kou"))
def test_fused_greeks_equal_bump_and_reprice(kind):
    """delta/gamma/rho read off one simulation == GreekMixin's re-pricings on CRN."""
    from optpricing_b200.techniques.base import GreekMixin
    model = MODELS[kind]()
    mk = lambda: ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=24, seed=8)  # noqa: E731
    for opt in (CALL, PUT):
        a, b = mk(), mk()
        assert a.delta(opt, STQ, model, RT) == pytest.approx(GreekMixin.delta(b, opt, STQ, model, RT), rel=1e-8)
        assert a.gamma(opt, STQ, model, RT) == pytest.approx(GreekMixin.gamma(b, opt, STQ, model, RT), rel=1e-5)
        assert a.rho(opt, STQ, model, RT) == pytest.approx(GreekMixin.rho(b, opt, STQ, model, RT), rel=1e-8)
        # neither advanced the stream
        assert a.price(opt, STQ, model, RT).price == b.price(opt, STQ, model, RT).price


@pytest.mark.parametrize("cv", (False, True))
@pytest.mark.parametrize("kind", ("bsm", "merton", "kou"))
def test_one_simulation_vega_equals_bump_and_reprice(kind, cv):
    """vega from the stored (W, J) of ONE simulation == GreekMixin.vega's two re-pricings on
    common random numbers (greek_mixin.py:125-175), with and without the control variate."""
    from optpricing_b200.techniques.base import GreekMixin
    model = MODELS[kind]()
    mk = lambda: ob.MonteCarloTechnique(n_paths=(1 << 18) + 2, n_steps=25, seed=8,  # noqa: E731
                                        control_variate=cv)
    for opt in (CALL, PUT):
        a, b = mk(), mk()
        launches0 = _engine.get_engine().launch_count()
        got = a.vega(opt, STQ, model, RT)
        assert _engine.get_engine().launch_count() - launches0 == 4     # 1 simulation + 1 sweep
        assert got == pytest.approx(GreekMixin.vega(b, opt, STQ, model, RT), rel=1e-6)
        assert a.price(opt, STQ, model, RT).price == b.price(opt, STQ, model, RT).price


@pytest.mark.parametrize("kind", ("bsm", "heston", "kou"))
def test_fused_greeks_with_control_variate(kind):
    from optpricing_b200.techniques.base import GreekMixin
    model = MODELS[kind]()
    mk = lambda: ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=24, seed=8,  # noqa: E731
                                        control_variate=True)
    a, b = mk(), mk()
    assert a.delta(PUT, STQ, model, RT) == pytest.approx(GreekMixin.delta(b, PUT, STQ, model, RT), rel=1e-7)
    assert a.rho(CALL, STQ, model, RT) == pytest.approx(GreekMixin.rho(b, CALL, STQ, model, RT), rel=1e-7)


@pytest.mark.parametrize("cv", (False, True))
@pytest.mark.parametrize("anti", (True, False))
@pytest.mark.parametrize("kind", ("sabr", "sabr_jump"))
def test_sabr_spot_scenarios_equal_bump_and_reprice(kind, anti, cv, jump_schedule):
    """SABR kinds: bumped-spot prices carried side by side through ONE simulation
    (k_european_scen) equal GreekMixin's re-pricings on common random numbers."""
    import dataclasses
    from optpricing_b200.techniques.base import GreekMixin
    if kind == "sabr" and jump_schedule == 0:
        pytest.skip("no jumps")
    model = MODELS[kind]()
    mk = lambda: ob.MonteCarloTechnique(n_paths=200_001, n_steps=37, seed=21, antithetic=anti,  # noqa: E731
                                        control_variate=cv)
    a, b = mk(), mk()
    fs = (1.001, 1.0, 0.999)
    launches0 = _engine.get_engine().launch_count()
    got = a._scenario_prices(PUT, STQ, model, RT, fs)
    assert _engine.get_engine().launch_count() - launches0 == 2        # one simulation + finish
    want = GreekMixin._reprice(b, [(PUT, dataclasses.replace(STQ, spot=STQ.spot * f), model, RT)
                                   for f in fs])
    np.testing.assert_allclose(got, want, rtol=1e-9)
    assert a.delta(CALL, STQ, model, RT) == pytest.approx(GreekMixin.delta(b, CALL, STQ, model, RT), rel=1e-6)
    assert a.gamma(CALL, STQ, model, RT) == pytest.approx(GreekMixin.gamma(b, CALL, STQ, model, RT),
                                                          rel=1e-4, abs=1e-7)
    assert np.isnan(a.vega(CALL, STQ, model, RT))                       # greek_mixin.py:156-157
    assert a.price(CALL, STQ, model, RT).price == b.price(CALL, STQ, model, RT).price


@pytest.mark.parametrize("cv", (False, True))
@pytest.mark.parametrize("kind", ("bsm", "heston", "kou", "bates", "sabr_jump"))
def test_price_with_greeks_is_one_simulation_and_equals_the_mixin(kind, cv):
    from optpricing_b200.techniques.base import GreekMixin
    model = MODELS[kind]()
    mk = lambda: ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=24, seed=5, control_variate=cv)  # noqa: E731
    a, b = mk(), mk()
    eng = _engine.get_engine()
    sims0 = eng.n_simulations
    res = a.price_with_greeks(CALL, STQ, model, RT, greeks=("delta", "gamma", "vega", "rho")
                              if kind != "sabr_jump" else ("delta", "gamma", "vega"))
    assert eng.n_simulations - sims0 == 1
    assert isinstance(res, ob.PricingResult)
    # the mixin's estimators on the same stream position
    want_delta = GreekMixin.delta(b, CALL, STQ, model, RT)
    want_gamma = GreekMixin.gamma(b, CALL, STQ, model, RT)
    assert res.greeks["delta"] == pytest.approx(want_delta, rel=1e-6)
    assert res.greeks["gamma"] == pytest.approx(want_gamma, rel=1e-4, abs=1e-7)
    if kind != "sabr_jump":
        assert res.greeks["rho"] == pytest.approx(GreekMixin.rho(b, CALL, STQ, model, RT), rel=1e-6)
    if "sigma" in model.params:
        assert res.greeks["vega"] == pytest.approx(GreekMixin.vega(b, CALL, STQ, model, RT), rel=1e-6)
    else:
        assert np.isnan(res.greeks["vega"])
    # same seed as one price() (the sweep adds the payoffs in another order: last-bit differences)
    assert res.price == pytest.approx(b.price(CALL, STQ, model, RT).price, rel=1e-12)
    # both generators advanced once
    assert a.rng.bit_generator.state == b.rng.bit_generator.state


@pytest.mark.parametrize("cv", (False, True))
@pytest.mark.parametrize("anti", (True, False))
def test_bsm_theta_from_one_simulation_equals_the_mixin(anti, cv):
    """GBM theta: both maturity-bumped prices read off the Brownian sums of ONE simulation
    (a (drift, sigma) scenario of the vega sweep) equal GreekMixin.theta's two re-pricings on
    common random numbers; price_with_greeks delivers it from the same single simulation; a
    jump model falls back to the re-pricings."""
    from optpricing_b200.techniques.base import GreekMixin
    model = ob.BSMModel({"sigma": 0.25})
    mk = lambda **k: ob.MonteCarloTechnique(n_paths=1 << 18, n_steps=24, seed=5, antithetic=anti,  # noqa: E731
                                            control_variate=cv, **k)
    a, b = mk(), mk()
    eng = _engine.get_engine()
    sims0 = eng.n_simulations
    got = a.theta(PUT, STQ, model, RT)
    assert eng.n_simulations - sims0 == 1
    want = GreekMixin.theta(b, PUT, STQ, model, RT)
    assert got == pytest.approx(want, rel=1e-5)
    assert a.rng.bit_generator.state == b.rng.bit_generator.state      # greeks do not advance it
    sims0 = eng.n_simulations
    res = a.price_with_greeks(CALL, STQ, model, RT, greeks=("delta", "vega", "theta"))
    assert eng.n_simulations - sims0 == 1
    assert res.greeks["theta"] == pytest.approx(GreekMixin.theta(b, CALL, STQ, model, RT), rel=1e-5)
    # the time grid follows the maturity: same step count for T +- h, still one simulation
    c, d = mk(steps_per_year=24), mk(steps_per_year=24)
    assert c.theta(PUT, STQ, model, RT) == pytest.approx(GreekMixin.theta(d, PUT, STQ, model, RT), rel=1e-5)
    # jump kinds re-price (their jump counts depend on lambda T)
    e, f = mk(), mk()
    sims0 = eng.n_simulations
    t_m = e.theta(PUT, STQ, ob.MertonJumpModel(), RT)
    assert eng.n_simulations - sims0 == 2
    assert t_m == pytest.approx(GreekMixin.theta(f, PUT, STQ, ob.MertonJumpModel(), RT), rel=1e-9)


# ------------------------------------------------------------------ BASELINE.json configs at full size
def test_config1_bsm_100k_x_252(price_goldens):
    """C1: BSM European call, 100k paths x 252 steps, antithetic (the reference's CPU case:
    its own run gave 10.4406, closed form 10.4506)."""
    t = ob.MonteCarloTechnique(n_paths=100_000, n_steps=252, seed=42)
    got = t.price(CALL, ST, ob.BSMModel({"sigma": 0.2}), RT).price
    assert abs(got - price_goldens["G1_closed_form"]) < 3 * t.last_stats["stderr"] + 1e-12
    assert t.last_stats["n_samples"] == 50_000


def test_config2_heston_2p24_x_252_matches_reference_statistics():
    """C2: within 3 sigma of the reference's price.  The reference cannot run 2^24 paths;
    its estimator at 400k paths x 252 gave 10.02-10.03 (SE 0.009) and the
    oracle port at 2^20 paths pins the mean more tightly."""
    t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=252, seed=42)
    got = t.price(CALL, ST, ob.HestonModel(), RT).price
    want, se_ref = _oracle_price_and_se("heston", ob.HestonModel().params, 100.0, 100.0, 1.0, 0.05,
                                        0.0, True, 1 << 19, 252, 3)
    assert abs(got - want) < 3 * math.hypot(t.last_stats["stderr"], se_ref)
    assert t.last_stats["stderr"] < 0.0016


def test_config3_bates_strike_maturity_sweep(price_goldens):
    """C3: 64 strikes x 4 maturities, 2^22 paths per maturity, n_steps = 252 T."""
    model = ob.BatesModel()
    strikes = np.linspace(70, 130, 64)
    total = 0
    for T in (0.25, 0.5, 1.0, 2.0):
        t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=int(252 * T), seed=7,
                                   correlation="independent")
        opts = [ob.Option(float(k), T, ob.OptionType.CALL) for k in strikes]
        p = t.price_many(opts, ST, model, RT)
        total += p.size
        assert np.all(np.diff(p) < 0)                       # calls decrease in strike
        intrinsic = np.maximum(100.0 - strikes * math.exp(-0.05 * T), 0)
        assert np.all(p > intrinsic - 4 * 0.02)             # no-arbitrage lower bound
        if T == 1.0:                                        # vs the reference's FFT (K=100), Euler bias allowed
            k100 = np.interp(100.0, strikes, p)
            assert abs(k100 - price_goldens["bates_fft"]) < 0.08
    assert total == 256


def test_config4_american_put_2p22_x_50_cubic(price_goldens):
    """C4: American put, GBM, 2^22 paths x 50 dates, cubic basis; truth = CRR-1001 12.2777
    (LSMC at 50 dates with a cubic basis is biased low by ~0.02-0.03, as is the reference)."""
    aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
    t = ob.AmericanMonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=42, lsm_degree=3)
    got = t.price(aput, STQ, ob.BSMModel({"sigma": 0.2}), RT).price
    assert got == pytest.approx(price_goldens["crr1001_american_put"], abs=0.05)
    assert got == pytest.approx(price_goldens["G6_lsmc_bsm_50000x50_seed42_deg3"], abs=0.05)


@pytest.mark.parametrize("kind", ("sabr_jump", "kou"))
def test_config5_control_variates_and_bumps_2p24_x_365(kind):
    """C5: SABR-with-jumps and Kou, 2^24 paths x 365 steps, control variate, delta / vega."""
    model = MODELS[kind]()
    t = ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=365, seed=11, control_variate=True)
    p = t.price(CALL, STQ, model, RT).price
    s = t.last_stats
    assert s["n_samples"] == 1 << 23
    assert abs(s["control_mean"] - s["control_expected"]) < 4 * s["control_stderr"]
    assert s["cv_stderr"] < s["stderr"] and abs(s["cv_price"] - s["price"]) < 4 * s["stderr"]
    assert p == s["cv_price"]
    t2 = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=365, seed=11)
    d = t2.delta(CALL, STQ, model, RT, h_frac=1e-3)
    assert 0.5 < d < 0.8
    v = t2.vega(CALL, STQ, model, RT)
    if kind == "kou":
        assert 20 < v < 45                                   # has 'sigma'
    else:
        assert np.isnan(v)                                   # greek_mixin.py:156-157


# ------------------------------------------------------------------ single-step models
LEVY_CASES = ("vg", "vg_small_shape", "nig", "cgmy", "hyperbolic", "cev", "cev_gamma1")
LEVY_MODELS = {"vg": ob.VarianceGammaModel, "nig": ob.NIGModel, "cgmy": ob.CGMYModel,
               "hyperbolic": ob.HyperbolicModel, "cev": ob.CEVModel}


@pytest.mark.parametrize("case", LEVY_CASES)
def test_single_step_device_samplers_follow_the_reference_law(levy_goldens, case):
    """Device gamma / inverse-Gaussian / GIG / non-central chi-square samplers against
    2M draws of the reference's numpy/scipy samplers: price within 4 combined standard
    errors, empirical CDF at the reference's quantiles within 5 sigma."""
    e = levy_goldens[case]
    model = LEVY_MODELS[case.split("_")[0]](e["params"])
    call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
    t = ob.MonteCarloTechnique(n_paths=1 << 22, seed=5)
    price = t.price(call, STQ, model, RT).price
    se = math.hypot(t.last_stats["stderr"], e["call_2m_stderr"])
    assert abs(price - e["call_2m_price"]) < 4 * se, (price, e["call_2m_price"], se)
    spots = t._device_terminal(model, 100.0, 0.05, 0.01, 1.0)
    assert spots.shape == (1 << 22,) and np.all(np.isfinite(spots)) and np.all(spots > 0)
    assert spots.mean() == pytest.approx(e["ST_2m_mean"], abs=5 * spots.std() * math.sqrt(1 / spots.size + 1 / 2e6))
    for q, x in zip(levy_goldens["_quantile_levels"], e["ST_2m_quantiles"]):
        sd = math.sqrt(q * (1 - q) * (1 / spots.size + 1 / 2e6))
        assert abs(float((spots <= x).mean()) - q) < 5 * sd, (case, q)


@pytest.mark.parametrize("kind", ("vg", "hyperbolic", "cev"))
def test_single_step_price_within_3_sigma_of_oracle(kind):
    model = LEVY_MODELS[kind]()
    put = ob.Option(strike=95.0, maturity=0.5, option_type=ob.OptionType.PUT)
    t = ob.MonteCarloTechnique(n_paths=1 << 21, seed=3)
    got = t.price(put, STQ, model, RT).price
    np.random.seed(17)
    want, STv = orc.price_single_step(kind, model.params, 100.0, 95.0, 0.5, 0.05, 0.01, False,
                                      400_000, np.random.default_rng(8), return_terminal=True)
    se_o = math.exp(-0.025) * np.maximum(95.0 - STv, 0).std(ddof=1) / math.sqrt(STv.size)
    assert abs(got - want) < 3 * math.hypot(t.last_stats["stderr"], se_o)


def test_single_step_samples_do_not_depend_on_the_partition(eng):
    levy = _engine.make_levy("vg", ob.VarianceGammaModel().params, 0.05, 1.0, 100.0)
    torch = eng.torch
    mom = torch.zeros(6, dtype=torch.float64, device=eng.device)
    whole = torch.empty(5000, dtype=torch.float64, device=eng.device)
    eng.levy_terminal_async(levy, 100.0, 0.01, 77, 0, 5000, [100.0], [1], mom, whole)
    parts = []
    for off, cnt in ((0, 1234), (1234, 3766)):
        part = torch.empty(cnt, dtype=torch.float64, device=eng.device)
        eng.levy_terminal_async(levy, 100.0, 0.01, 77, off, cnt, [100.0], [1], mom, part)
        parts.append(part)
    assert torch.equal(whole, torch.cat(parts))


def test_single_step_errors_are_the_references():
    call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
    with pytest.raises(NotImplementedError, match="only implemented for Y=1"):
        ob.MonteCarloTechnique(n_paths=100).price(call, STQ, ob.CGMYModel(), RT)    # default Y = 1.2


def test_cev_sample_terminal_spot_is_the_device_sampler(levy_goldens):
    """models/cev.py:49-96 signature; reproducible through numpy's global seed as in the reference."""
    m = ob.CEVModel()
    np.random.seed(5)
    a = m.sample_terminal_spot(100.0, 0.05, 1.0, 100_000)
    np.random.seed(5)
    b = m.sample_terminal_spot(100.0, 0.05, 1.0, 100_000)
    assert a.shape == (100_000,) and np.array_equal(a, b)
    e = levy_goldens["cev"]
    assert a.mean() == pytest.approx(e["ST_2m_mean"], abs=5 * a.std() * math.sqrt(1 / a.size + 1 / 2e6))
    # a user-supplied sampler is called as is (monte_carlo.py:103-104)
    m.sample_terminal_spot = lambda S0, r, T, size: np.full(size, 120.0)
    call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
    assert ob.MonteCarloTechnique(n_paths=10).price(call, STQ, m, RT).price == pytest.approx(
        20.0 * math.exp(-0.05))


# ------------------------------------------------------------------ Dupire local volatility
@pytest.mark.parametrize("surface", ("flat", "term"))
def test_dupire_fed_kernel_matches_reference_body_when_the_table_is_exact(dupire_vectors, surface):
    """Surfaces constant in S are represented exactly by the table: <= 1e-12 against the
    reference's own (un-jitted) dupire_kernel on the same increments."""
    z, meta = dupire_vectors
    got = mc_kernels.dupire_kernel(meta["n"], meta["m"], meta["log_s0"], meta["r"], meta["q"],
                                   meta["dt"], z["dw"], orc.dupire_surfaces()[surface])
    np.testing.assert_allclose(got, z[surface], rtol=RTOL, atol=0)


def test_dupire_fed_kernel_skew_surface(dupire_vectors):
    """A surface curved in log S: the device step equals the numpy restatement of its own
    interpolation rule to 1e-12, and the reference with the exact callable to the
    interpolation error of the default 1024-node grid."""
    from optpricing_b200._engine import sample_local_vol
    z, meta = dupire_vectors
    f = orc.dupire_surfaces()["skew"]
    args = (meta["n"], meta["m"], meta["log_s0"], meta["r"], meta["q"], meta["dt"], z["dw"])
    got = mc_kernels.dupire_kernel(*args, f)
    table, lo, h = sample_local_vol(f, meta["m"], meta["dt"], meta["log_s0"], meta["r"], meta["q"])
    want = orc.dupire_kernel(*args, orc.table_surface(table, lo, h, meta["dt"]))
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    np.testing.assert_allclose(got, z["skew"], rtol=0, atol=2e-6)
    fine = mc_kernels.dupire_kernel(*args, f, n_s=16384)
    np.testing.assert_allclose(fine, z["skew"], rtol=0, atol=1e-8)


def test_dupire_philox_prices(price_goldens):
    call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
    # a flat surface is Black-Scholes: closed form within 3 standard errors
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=12)
    flat = ob.DupireLocalVolModel({"vol_surface": lambda T, K: 0.2})
    p = t.price(call, ST, flat, RT).price
    assert abs(p - price_goldens["G1_closed_form"]) < 3 * t.last_stats["stderr"] + 2e-3
    # a skewed surface: within 3 combined standard errors of the oracle on its own draws
    f = orc.dupire_surfaces()["skew"]
    t = ob.MonteCarloTechnique(n_paths=1 << 21, n_steps=40, seed=3)
    got = t.price(call, STQ, ob.DupireLocalVolModel({"vol_surface": f}), RT).price
    rng = np.random.default_rng(5)
    n, m, dt = 200_000, 40, 1.0 / 40
    dw = rng.standard_normal((n, m)) * math.sqrt(dt)
    a = np.exp(orc.dupire_kernel(n, m, math.log(100.0), 0.05, 0.01, dt, dw, f))
    b = np.exp(orc.dupire_kernel(n, m, math.log(100.0), 0.05, 0.01, dt, -dw, f))
    pair = 0.5 * (np.maximum(a - 100.0, 0) + np.maximum(b - 100.0, 0)) * math.exp(-0.05)
    se = math.hypot(t.last_stats["stderr"], pair.std(ddof=1) / math.sqrt(n))
    assert abs(got - pair.mean()) < 3 * se, (got, pair.mean(), se)
    # the scheme is a martingale step by step
    s = t.last_stats
    assert abs(s["control_mean"] - s["control_expected"]) < 4 * s["control_stderr"]
    # numpy mode runs the fed kernel on the reference's host draws
    # -- WITHOUT the antithetic pass, as the reference does for local-vol models
    # (monte_carlo.py:247): n_paths / 2 plain samples, i.e. what BSM prices from the same
    # 10 000 draws with antithetic=False
    tn = ob.MonteCarloTechnique(n_paths=20000, n_steps=10, seed=0, rng_mode="numpy")
    tb = ob.MonteCarloTechnique(n_paths=10000, n_steps=10, seed=0, rng_mode="numpy", antithetic=False)
    assert tn.price(call, ST, flat, RT).price == pytest.approx(
        tb.price(call, ST, ob.BSMModel({"sigma": 0.2}), RT).price, rel=1e-12)
    assert tn.last_stats["n_samples"] == 10000
    with pytest.raises(TypeError, match="requires an SDE model"):
        ob.AmericanMonteCarloTechnique(n_paths=100, n_steps=5).price(call, ST, flat, RT)


# ------------------------------------------------------------------ two ranks on one box
def test_two_ranks_match_one_gpu_nvlink_and_nccl_exchange():
    """torchrun x 2 over tests/dist_gpu_check.py: European moments all-reduced by NCCL, the
    LSMC Gram exchange both over NVLink peer memory inside the persistent kernel and as one
    NCCL all-reduce per date; each must equal the single-GPU price to rounding."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29533", os.path.join(root, "tests", "dist_gpu_check.py")],
                       env={**os.environ, "DIST_CHECK_QUICK": "1"}, capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0 and "dist_gpu_check: PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ------------------------------------------------------------------ optional fp32 mode
@pytest.mark.parametrize("kind", KINDS)
def test_fp32_mode_agrees_with_fp64_statistically(kind):
    """precision="fp32": different arithmetic and a different Philox stream, the same law --
    prices within 3 combined standard errors (+ 2e-3 for single-precision drift) of fp64,
    and the scheme's martingale mean within 4 sigma."""
    model = MODELS[kind]()
    put = ob.Option(strike=105.0, maturity=1.0, option_type=ob.OptionType.PUT)
    kw = dict(n_paths=1 << 22, n_steps=64, seed=21)
    t64 = ob.MonteCarloTechnique(**kw)
    t32 = ob.MonteCarloTechnique(precision="fp32", **kw)
    p64 = t64.price(put, STQ, model, RT).price
    p32 = t32.price(put, STQ, model, RT).price
    se = math.hypot(t64.last_stats["stderr"], t32.last_stats["stderr"])
    assert abs(p32 - p64) < 3 * se + 2e-3, (kind, p32, p64, se)
    s = t32.last_stats
    assert abs(s["control_mean"] - s["control_expected"]) < 4 * s["control_stderr"] + 1e-3, kind
    assert t32.price(put, STQ, model, RT).price != p32          # the stream advances as in fp64
    # the strike sweep on stored fp32-mode terminals reproduces the single-contract price
    many = ob.MonteCarloTechnique(precision="fp32", **kw).price_many([put, CALL], STQ, model, RT)
    assert many[0] == pytest.approx(p32, rel=1e-10)


def test_fp32_mode_restrictions():
    with pytest.raises(ValueError, match="device RNG only"):
        ob.MonteCarloTechnique(precision="fp32", rng_mode="numpy")
    with pytest.raises(ValueError, match="precision must be"):
        ob.MonteCarloTechnique(precision="fp16")
    flat = ob.DupireLocalVolModel()
    with pytest.raises(_engine.EngineError, match="not available for Dupire"):
        ob.MonteCarloTechnique(precision="fp32", n_paths=1000, n_steps=4).price(CALL, ST, flat, RT)


def test_implied_volatility_from_one_simulation(price_goldens):
    """One simulation, every trial sigma a sweep over its Brownian sums: recovers the vol of a
    Black-Scholes target to the Monte Carlo error, deterministically, with 2 + 2k launches."""
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=4, seed=6)
    eng = _engine.get_engine()
    sims0 = eng.n_simulations
    iv = t.implied_volatility(CALL, ST, ob.BSMModel(), RT, target_price=price_goldens["G1_closed_form"])
    assert eng.n_simulations - sims0 == 1
    assert iv == pytest.approx(0.2, abs=3e-4)
    t2 = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=4, seed=6)
    assert t2.implied_volatility(CALL, ST, ob.BSMModel(), RT,
                                 target_price=price_goldens["G1_closed_form"]) == iv
    # an unreachable target has no root: NaN as in the reference (iv_mixin.py:81-116)
    assert math.isnan(t.implied_volatility(CALL, ST, ob.BSMModel(), RT, target_price=1e9))


def test_steps_per_year_follows_the_maturity():
    """steps_per_year=48: a 0.5-year option is priced on 24 steps, a 2-year one on 96, and one
    price_many call over both maturities equals separate fixed-grid techniques (same seeds)."""
    model = ob.HestonModel()
    o_short = ob.Option(strike=100.0, maturity=0.5, option_type=ob.OptionType.CALL)
    o_long = ob.Option(strike=95.0, maturity=2.0, option_type=ob.OptionType.PUT)
    a = ob.MonteCarloTechnique(n_paths=1 << 16, steps_per_year=48, seed=4)
    b = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=24, seed=4)
    assert a.price(o_short, STQ, model, RT).price == b.price(o_short, STQ, model, RT).price
    assert a.n_steps == 100                                   # the ctor default is untouched
    many = ob.MonteCarloTechnique(n_paths=1 << 16, steps_per_year=48, seed=4).price_many(
        [o_short, o_long], STQ, model, RT)
    s1 = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=24, seed=4)
    want_short = s1.price_many([o_short], STQ, model, RT)[0]
    # the second maturity draws the technique's second seed: replay it on a 96-step grid
    s2 = ob.MonteCarloTechnique(n_paths=1 << 16, n_steps=96, seed=4)
    s2.rng.integers(0, 2**63 - 1, dtype=np.int64)
    want_long = s2.price_many([o_long], STQ, model, RT)[0]
    assert many[0] == pytest.approx(want_short, rel=1e-12)
    assert many[1] == pytest.approx(want_long, rel=1e-12)


def test_report_stats_travel_with_the_price(price_goldens):
    t = ob.MonteCarloTechnique(n_paths=1 << 20, n_steps=8, seed=2, report_stats=True)
    res = t.price(CALL, ST, ob.BSMModel({"sigma": 0.2}), RT)
    g = res.greeks
    assert g["n_samples"] == (1 << 19) and 0 < g["stderr"] < 0.02
    assert g["ci95_low"] < res.price < g["ci95_high"]
    assert abs(res.price - price_goldens["G1_closed_form"]) < 4 * g["stderr"]
    assert ob.MonteCarloTechnique(n_paths=1 << 12, n_steps=4, seed=2).price(
        CALL, ST, ob.BSMModel(), RT).greeks == {}                     # default: the reference's result


def test_pricings_from_two_threads_do_not_share_buffers():
    """Engine.buffer() hands out cached device buffers; the techniques' entry points run under
    one re-entrant process-wide lock, so concurrent callers get what serial callers get."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = [("heston", 1 << 20, 40, 3), ("bsm", 1 << 21, 16, 4), ("kou", 1 << 19, 64, 5),
            ("bates", 1 << 20, 24, 6)] * 3

    def price(job):
        kind, n, m, seed = job
        t = ob.MonteCarloTechnique(n_paths=n, n_steps=m, seed=seed)
        return (t.price(CALL, STQ, MODELS[kind](), RT).price,
                t.delta(CALL, STQ, MODELS[kind](), RT))
    serial = [price(j) for j in jobs]
    with ThreadPoolExecutor(max_workers=6) as ex:
        threaded = list(ex.map(price, jobs))
    assert threaded == serial


def test_large_jump_intensity_whole_path_draw(eng):
    """lambda T far beyond where exp(-lambda T) is usable (the whole-path count is drawn in
    chunks of intensity <= 16): the martingale control mean and the one-step exact law hold."""
    p = {"sigma": 0.1, "lambda": 900.0, "mu_j": -0.001, "sigma_j": 0.004}
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=1, seed=3, control_variate=True)
    t.price(CALL, STQ, ob.MertonJumpModel(p), RT)
    s = t.last_stats
    assert abs(s["control_mean"] - s["control_expected"]) < 4 * s["control_stderr"]
    # one step: log S_T is N(m, s^2) mixed over N ~ Poisson(900); its variance is
    # sigma^2 T + lambda T (mu_j^2 + sigma_j^2)
    term = t._simulate_sde_path(ob.MertonJumpModel(p), 100.0, 0.05, 0.01, 1.0)
    want_var = 0.1 ** 2 + 900.0 * (0.001 ** 2 + 0.004 ** 2)
    assert np.log(term).var() == pytest.approx(want_var, rel=5 * math.sqrt(2 / term.size) * 2)


def test_fed_jump_buffer_overrun_is_contained(eng):
    """Counts asking for more jump sizes than the buffer holds: NaN paths, no out-of-bounds read."""
    n, m = 64, 8
    dw = np.zeros((n, m))
    counts = np.zeros((n, m), dtype=np.int64)
    counts[:, 0] = 1                               # 64 sizes wanted, consumed in path order ...
    mdl = _engine.make_model("merton", ob.MertonJumpModel().params, 0.05, 0.0)
    term, _ = eng.fed("merton", mdl, math.log(100.0), 1.0 / m, dw, None, counts,
                      np.full(40, 0.01), sign=1.0, path_sum_mode=0, want_terminal=True,
                      want_paths=False)          # ... 40 given
    assert np.isfinite(term[:40]).all() and np.isnan(term[40:]).all()
