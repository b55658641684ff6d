"""
Path-by-path pinning of the THROUGHPUT kernels -- the ones bench.py times.

The <= 1e-12 parity tests of test_gpu_parity.py run the fed-draw kernels (the reference's own
increments in, the literal step).  The device-RNG kernels (k_european, k_european_wide,
k_lsmc_paths) generate their normals in registers and re-associate the step (deferred drift,
one square root per path of v+ * (-2 ln u), log-vol SABR, jumps drawn once at the end), so the
same-law argument has to be checked where it can be checked exactly:

1. the device's word -> normal mapping against a numpy restatement (tests/philox_ref.py);
2. REPLAY: dump the device normals of every (draw, pair) with optmc_normals, build the
   increments the way the kernel mixes them, run the ORACLE's literal recurrence on them and
   compare with the terminal values / path store the throughput kernel wrote for the same
   seed -- path by path, <= 1e-9 relative (the 1.3e-12 root and the re-association bound it).
   Jump kinds: the diffusion part is replayed, the jump part is the path's summed log-jumps,
   which the kernel hands out through the aux buffer;
3. statistics at a resolution that sees a 0.1 % bias: 2^26 device paths against 2^22 oracle
   paths x 252 steps for all seven kinds (sigma ~ 0.005 on a price of ~10), and 2^30 paths of
   the exact one-step laws against closed forms (sigma ~ 3e-4).

Everything goes through the C ABI (ctypes); the oracle is only the checker.
"""
import json
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch  # noqa: E402

import optpricing_b200 as ob  # noqa: E402
from optpricing_b200 import _engine  # noqa: E402
from oracle import reference_mc as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
S0, K, T, R, Q = 100.0, 100.0, 1.0, 0.05, 0.01
MODELS = {"bsm": ob.BSMModel, "heston": ob.HestonModel, "merton": ob.MertonJumpModel,
          "bates": ob.BatesModel, "kou": ob.KouModel, "sabr": ob.SABRModel,
          "sabr_jump": ob.SABRJumpModel}
TWO_FACTOR = ("heston", "bates", "sabr", "sabr_jump")


@pytest.fixture(scope="module")
def eng():
    return _engine.get_engine()


# ------------------------------------------------------------------ 1. the mapping
def test_device_normals_equal_the_numpy_restatement_of_the_mapping(eng):
    from philox_ref import normals_from_words, pair_words
    seed, n = 0x1234567890ABCDEF, 1 << 16
    for pair in (0, 1, 2, 7, 250):
        got = eng.normals(seed=seed, n=n, step=pair)
        wa, wb = pair_words(seed, np.arange(n), pair)
        a, b = normals_from_words(wa, wb)
        # table + degree-4 polynomial for the logarithm, one Newton step for the root
        np.testing.assert_allclose(got[0::2], a, rtol=0, atol=5e-11)
        np.testing.assert_allclose(got[1::2], b, rtol=0, atol=5e-11)
    eng.set_int("normals_impl", 0)          # the same mapping through the CUDA math library
    try:
        lib = eng.normals(seed=seed, n=n, step=7)
    finally:
        eng.set_int("normals_impl", 1)
    np.testing.assert_allclose(lib, eng.normals(seed=seed, n=n, step=7), rtol=0, atol=5e-11)


def test_angle_lattice_and_radius_grid(eng):
    """2^18 equally spaced directions, radius from a 46-bit uniform: |z| < 8.08."""
    z = eng.normals(seed=99, n=1 << 20, step=0)
    a, b = z[0::2], z[1::2]
    ang = np.arctan2(b, a) / (2 * np.pi) * 2 ** 18 - 0.5
    assert np.abs(ang - np.round(ang)).max() < 1e-4
    assert np.hypot(a, b).max() < math.sqrt(2 * 47 * math.log(2)) + 1e-9


# ------------------------------------------------------------------ 2. replay
def _device_pairs(eng, seed, n_draws, n_pairs):
    a = np.empty((n_draws, n_pairs))
    b = np.empty((n_draws, n_pairs))
    for p in range(n_pairs):
        z = eng.normals(seed=seed, n=n_draws, step=p)
        a[:, p], b[:, p] = z[0::2], z[1::2]
    return a, b


def _increments(kind, params, a, b, n_steps, dt, correlation):
    """What the step loop feeds the model from the unit normals (a, b) of each pair."""
    sd = math.sqrt(dt)
    if kind in TWO_FACTOR:
        if correlation == _engine.CORR_REFERENCE_EUROPEAN:       # monte_carlo.py:222-231
            rho = params["rho"]
            dw_v, dw_u = sd * a, sd * b
            return {"dw1": rho * dw_v + math.sqrt(1 - rho * rho) * dw_u, "dw2": dw_v}
        return {"dw1": sd * a, "dw2": sd * b}                     # american_monte_carlo.py:181-185
    dw = np.empty((a.shape[0], 2 * a.shape[1]))
    dw[:, 0::2], dw[:, 1::2] = sd * a, sd * b                     # a pair feeds two steps
    return {"dw": np.ascontiguousarray(dw[:, :n_steps])}


def _oracle_pass(kind, params, n_draws, n_steps, dt, inc, sign, want_paths=False):
    sp = {"n_paths": n_draws, "n_steps": n_steps, **orc.kernel_params(kind, params, R, Q, dt)}
    sp["s0" if kind in orc.SABR_KINDS else "log_s0"] = S0 if kind in orc.SABR_KINDS else math.log(S0)
    for k, v in inc.items():
        sp[k] = sign * v
    if kind in orc.JUMP_KINDS:                    # diffusion + compensated drift only
        sp["jump_counts"] = np.zeros((n_draws, n_steps), dtype=np.int64)
    return orc.run_kernel(kind, sp, jump_buf=np.zeros(0), want_paths=want_paths)


REPLAY_CASES = [  # (kind, n_draws, n_steps): 3001 draws -> compact kernel, 160000 -> one block per SM
    ("bsm", 3001, 37), ("bsm", 160000, 64), ("bsm", 160000, 37),
    ("heston", 3001, 37), ("heston", 160000, 37), ("heston", 160000, 252),
    ("sabr", 3001, 37), ("sabr", 160000, 37),
    ("merton", 3001, 37), ("merton", 160000, 37),
    ("kou", 3001, 11), ("kou", 160000, 37),
    ("bates", 3001, 37), ("bates", 160000, 37),
    ("sabr_jump", 3001, 37), ("sabr_jump", 160000, 37),
]


@pytest.mark.parametrize("kind,n_draws,n_steps", REPLAY_CASES)
def test_european_throughput_kernel_replays_through_the_oracle(eng, kind, n_draws, n_steps):
    seed = 20261020 + n_steps
    params = dict(MODELS[kind]().params)
    if kind == "sabr_jump":
        # its jumps act on S between the Euler steps, so a path is not (diffusion, sum of
        # jumps): replay the kernel's diffusion + compensator code path with the jumps off
        # (the jump law itself: section 3 and test_european_sabr_jump_schedule_...)
        params["lambda"] = 0.0
    dt = T / n_steps
    mdl = _engine.make_model(kind, params, R, Q)
    sim = eng.make_sim(S0, T, n_draws, n_steps, True, seed, _engine.CORR_REFERENCE_EUROPEAN)
    mom = torch.zeros(_engine.NMOM, dtype=torch.float64, device=eng.device)
    term = torch.zeros(2 * n_draws, dtype=torch.float64, device=eng.device)
    aux = None
    if kind in ("bsm", "merton", "kou", "bates"):
        aux = torch.zeros(3 * n_draws, dtype=torch.float64, device=eng.device)
    eng.european_async(mdl, sim, [K], [1], mom, terminal=term, aux=aux)
    got = term.cpu().numpy()
    n_pairs = n_steps if kind in TWO_FACTOR else (n_steps + 1) // 2
    a, b = _device_pairs(eng, seed, n_draws, n_pairs)
    inc = _increments(kind, params, a, b, n_steps, dt, _engine.CORR_REFERENCE_EUROPEAN)
    want = np.concatenate([_oracle_pass(kind, params, n_draws, n_steps, dt, inc, +1.0),
                           _oracle_pass(kind, params, n_draws, n_steps, dt, inc, -1.0)])
    if aux is not None:
        x = aux.cpu().numpy()
        if kind != "bates":                        # W: the path's Brownian sum
            np.testing.assert_allclose(x[:n_draws], inc["dw"].sum(axis=1), rtol=0, atol=1e-11)
        want = want + np.concatenate([x[n_draws:2 * n_draws], x[2 * n_draws:]])
        if kind in ("merton", "kou", "bates"):
            lam_T = params["lambda"] * T
            frac = float(np.mean(x[n_draws:2 * n_draws] != 0.0))     # P(N > 0) = 1 - exp(-lambda T)
            assert abs(frac - (1 - math.exp(-lam_T))) < 5 * math.sqrt(0.25 / n_draws)
            # the partner shares the count (monte_carlo.py:247-255), not the sizes
            assert np.array_equal(x[n_draws:2 * n_draws] != 0.0, x[2 * n_draws:] != 0.0)
            both = x[n_draws:2 * n_draws] != 0.0
            assert not np.any(x[n_draws:2 * n_draws][both] == x[2 * n_draws:][both])
    want = want if kind in orc.SABR_KINDS else np.exp(want)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)
    # and the moments are the moments of exactly these terminal values
    pay = np.maximum(got - K, 0.0)
    pair = 0.5 * (pay[:n_draws] + pay[n_draws:])
    m = mom.cpu().numpy()
    assert m[0] == n_draws
    assert m[1] == pytest.approx(pair.sum(), rel=1e-12)
    assert m[2] == pytest.approx((pair * pair).sum(), rel=1e-12)


def test_replay_holds_for_a_rank_shard(eng):
    """Draw ids are global: a shard [offset, offset + n_local) of a big job (which selects the
    one-block-per-SM kernel by the JOB's size) reproduces the same paths."""
    kind, n_draws, n_steps, seed = "heston", 1 << 20, 24, 77
    off, cnt = 123457, 5000
    params = MODELS[kind]().params
    mdl = _engine.make_model(kind, params, R, Q)
    sim = eng.make_sim(S0, T, n_draws, n_steps, True, seed, _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
    mom = torch.zeros(_engine.NMOM, dtype=torch.float64, device=eng.device)
    term = torch.zeros(2 * cnt, dtype=torch.float64, device=eng.device)
    eng.european_async(mdl, sim, [K], [1], mom, terminal=term)
    a, b = _device_pairs(eng, seed, off + cnt, n_steps)
    inc = _increments(kind, params, a[off:], b[off:], n_steps, T / n_steps,
                      _engine.CORR_REFERENCE_EUROPEAN)
    want = np.exp(np.concatenate([_oracle_pass(kind, params, cnt, n_steps, T / n_steps, inc, s)
                                  for s in (+1.0, -1.0)]))
    np.testing.assert_allclose(term.cpu().numpy(), want, rtol=1e-9, atol=0)


@pytest.mark.parametrize("kind,n_steps", [("bsm", 50), ("bsm", 7), ("heston", 50), ("sabr", 13)])
def test_american_path_store_replays_through_the_oracle(eng, kind, n_steps):
    """k_lsmc_paths (the fill of BASELINE config 4's path store): every stored spot of every
    path equals the oracle's path kernel on the device's normals (path_kernels.py:17-324,
    independent draws as american_monte_carlo.py:181-185)."""
    n_draws, seed = 20000, 4242
    params = MODELS[kind]().params
    dt = T / n_steps
    mdl = _engine.make_model(kind, params, R, Q)
    sim = eng.make_sim(S0, T, n_draws, n_steps, True, seed, _engine.CORR_INDEPENDENT_DRAWS)
    store, ld = eng.lsmc_store(2 * n_draws, n_steps)
    eng.lsmc_fill_paths(mdl, sim, store, ld)
    got = store.cpu().numpy()[: ld * n_steps].reshape(n_steps, ld)[:, : 2 * n_draws]
    n_pairs = n_steps if kind in TWO_FACTOR else (n_steps + 1) // 2
    a, b = _device_pairs(eng, seed, n_draws, n_pairs)
    inc = _increments(kind, params, a, b, n_steps, dt, _engine.CORR_INDEPENDENT_DRAWS)
    want = np.concatenate([_oracle_pass(kind, params, n_draws, n_steps, dt, inc, s, want_paths=True)
                           for s in (+1.0, -1.0)])            # (2 n_draws, n_steps + 1)
    np.testing.assert_allclose(got, want[:, 1:].T, rtol=1e-9, atol=0)


# ------------------------------------------------------------------ 3. statistics
def _oracle_moments(kind, params, n_paths, n_steps, seed, chunk=1 << 16):
    """Mean and standard error of the discounted antithetic-pair payoff over n_paths oracle
    paths (monte_carlo.py:65-257 restated), chunks priced on all host threads."""
    orc.lib()
    n_chunks = n_paths // chunk

    def work(i):
        _, ST = orc.price_european(kind, params, S0, K, T, R, Q, True, chunk, n_steps, True,
                                   np.random.default_rng([seed, i]),
                                   jump_rs=np.random.RandomState(seed * 1000 + i),
                                   return_terminal=True)
        pay = np.maximum(ST - K, 0.0)
        pair = 0.5 * (pay[: chunk // 2] + pay[chunk // 2:])
        return pair.sum(), (pair * pair).sum()
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 1, 32)) as ex:
        parts = list(ex.map(work, range(n_chunks)))
    n = n_chunks * (chunk // 2)
    s, ss = sum(p[0] for p in parts), sum(p[1] for p in parts)
    mean = s / n
    var = (ss / n - mean * mean) * n / (n - 1)
    disc = math.exp(-R * T)
    return disc * mean, disc * math.sqrt(var / n)


@pytest.fixture(params=[1, 0], ids=["jump_schedule", "poisson_per_step"])
def jump_sampler(request, eng):
    eng.set_int("jump_schedule", request.param)
    yield request.param
    eng.set_int("jump_schedule", 1)


_ORACLE_CACHE = {}


@pytest.mark.parametrize("kind", tuple(MODELS))
def test_philox_price_against_2p22_oracle_paths(eng, kind, jump_sampler):
    """2^26 device paths x 252 steps against 2^22 oracle paths x 252 steps on numpy draws:
    combined sigma ~ 0.005 (0.05 % of the price).  Both jump samplers for SABR-jump."""
    if jump_sampler == 0 and kind != "sabr_jump":
        pytest.skip("the per-step Poisson sampler only differs for SABR-jump here")
    n_steps = 252
    model = MODELS[kind]()
    if kind not in _ORACLE_CACHE:
        _ORACLE_CACHE[kind] = _oracle_moments(kind, model.params, 1 << 22, n_steps, seed=31)
    want, se_ref = _ORACLE_CACHE[kind]
    t = ob.MonteCarloTechnique(n_paths=1 << 26, n_steps=n_steps, seed=1234)
    got = t.price(ob.Option(K, T, ob.OptionType.CALL), ob.Stock(S0, dividend=Q), model,
                  ob.Rate(R)).price
    se = math.hypot(t.last_stats["stderr"], se_ref)
    assert se < 0.012
    assert abs(got - want) < 3 * se, (kind, got, want, se, (got - want) / se)


def _kou_exact(p):
    """Kou call by Gil-Pelaez quadrature of the characteristic function (q = 0)."""
    from scipy.integrate import quad
    sg, lam, pu, e1, e2 = p["sigma"], p["lambda"], p["p_up"], p["eta1"], p["eta2"]
    comp = lam * ((pu * e1 / (e1 - 1)) + ((1 - pu) * e2 / (e2 + 1)) - 1)

    def cf(u):
        drift = math.log(S0) + (R - 0.5 * sg ** 2 - comp) * T
        return np.exp(1j * u * drift - 0.5 * sg ** 2 * u ** 2 * T
                      + lam * T * (pu * e1 / (e1 - 1j * u) + (1 - pu) * e2 / (e2 + 1j * u) - 1))
    k = math.log(K)
    f1 = lambda u: np.real(np.exp(-1j * u * k) * cf(u - 1j) / (1j * u * cf(-1j)))   # noqa: E731
    f2 = lambda u: np.real(np.exp(-1j * u * k) * cf(u) / (1j * u))                  # noqa: E731
    P1 = 0.5 + quad(f1, 1e-12, 400, limit=2000, epsabs=1e-13, epsrel=1e-12)[0] / math.pi
    P2 = 0.5 + quad(f2, 1e-12, 400, limit=2000, epsabs=1e-13, epsrel=1e-12)[0] / math.pi
    return math.exp(-R * T) * (float(np.real(cf(-1j))) * P1 - K * P2)


def _deep(model, n_steps, want, chunks=4, **kw):
    call, stock, rate = ob.Option(K, T, ob.OptionType.CALL), ob.Stock(S0), ob.Rate(R)
    t = ob.MonteCarloTechnique(n_paths=1 << 28, n_steps=n_steps, seed=2026, **kw)
    s = ss = 0.0
    for _ in range(chunks):
        s += t.price(call, stock, model, rate).price
        ss += t.last_stats["stderr"] ** 2
    price, se = s / chunks, math.sqrt(ss) / chunks
    return (price - want) / se, se


@pytest.mark.parametrize("kind", ("bsm", "merton", "kou"))
def test_exact_one_step_laws_at_2p30_paths(kind):
    """n_steps = 1: the log-Euler scheme samples the exact terminal law, so the price must
    equal the closed form (the reference's own, tests/golden/price_goldens.json; Kou:
    quadrature) within its standard error of ~3e-4: a bias of 1e-4 of the price in the device
    normals, the Poisson draw or the jump sizes would show."""
    g = json.load(open(os.path.join(HERE, "golden", "price_goldens.json")))
    model = MODELS[kind]() if kind != "bsm" else ob.BSMModel({"sigma": 0.2})
    want = {"bsm": g["G1_closed_form"], "merton": g["merton_closed_form"]}.get(kind)
    if kind == "kou":
        want = _kou_exact(model.params)
    z, se = _deep(model, 1, want)
    assert se < 6e-4
    assert abs(z) < 3.5, (kind, z, se)


@pytest.mark.parametrize("corr", ("reference", "independent"))
def test_heston_with_constant_variance_is_black_scholes_at_2p30_paths(corr):
    """vol_of_vol -> 0, v0 = theta: the two-factor step (mixed direction tables, one root of
    v+ * (-2 ln u) per path) must reproduce Black-Scholes with sigma = sqrt(v0), any step count."""
    g = json.load(open(os.path.join(HERE, "golden", "price_goldens.json")))
    hes = ob.HestonModel({"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 1e-9})
    z, se = _deep(hes, 8, g["G1_closed_form"], correlation=corr)
    assert se < 6e-4
    assert abs(z) < 3.5, (corr, z, se)


# ------------------------------------------------------------------ configs 3 and 5, against the oracle
def test_config3_correlation_mode_the_bench_times():
    """C3 as bench.py runs it (the reference's double correlation): Bates strike
    sweep at T = 1 against oracle prices of the same strikes on numpy draws."""
    strikes = np.linspace(70, 130, 64)
    model = ob.BatesModel()
    t = ob.MonteCarloTechnique(n_paths=1 << 22, steps_per_year=252, seed=7)
    opts = [ob.Option(float(k), 1.0, ob.OptionType.CALL) for k in strikes]
    got = t.price_many(opts, ob.Stock(S0, dividend=Q), model, ob.Rate(R))
    n = 1 << 20
    _, ST = orc.price_european("bates", model.params, S0, K, 1.0, R, Q, True, n, 252, True,
                               np.random.default_rng(5), jump_rs=np.random.RandomState(5),
                               return_terminal=True)
    disc = math.exp(-R)
    for i in (0, 16, 32, 48, 63):
        pay = np.maximum(ST - strikes[i], 0.0)
        pair = 0.5 * (pay[: n // 2] + pay[n // 2:])
        want, se_ref = disc * pair.mean(), disc * pair.std(ddof=1) / math.sqrt(n // 2)
        se = math.hypot(se_ref, se_ref / 2)          # the sweep's own error is half (4x the paths)
        assert abs(got[i] - want) < 3.5 * se, (strikes[i], got[i], want, se)


def test_config5_kou_delta_and_vega_against_the_oracle_bump_and_reprice():
    """C5's greeks: Kou delta (h = 0.1 % of spot) and vega (h = 1e-4) from the device RNG
    against the oracle's bump-and-reprice on common numpy draws (greek_mixin.py:59-73,
    125-175), as z-scores.  The standard error of a CRN greek is estimated from the pathwise
    differences of the oracle's two bumped runs."""
    p = ob.KouModel().params
    n, m = 1 << 18, 365
    stock, rate, call = ob.Stock(S0, dividend=Q), ob.Rate(R), ob.Option(K, T, ob.OptionType.CALL)

    def bumped(fn_params, fn_spot):
        out = []
        for sgn in (+1, -1):
            pp, s0 = fn_params(sgn), fn_spot(sgn)
            _, ST = orc.price_european("kou", pp, s0, K, T, R, Q, True, n, m, True,
                                       np.random.default_rng(9), jump_rs=np.random.RandomState(9),
                                       return_terminal=True)
            pay = math.exp(-R * T) * np.maximum(ST - K, 0.0)
            out.append(0.5 * (pay[: n // 2] + pay[n // 2:]))
        return out
    h = 1e-3 * S0
    up, dn = bumped(lambda s: p, lambda s: S0 + s * h)
    d = (up - dn) / (2 * h)
    want_delta, se_delta = d.mean(), d.std(ddof=1) / math.sqrt(d.size)
    hv = 1e-4
    up, dn = bumped(lambda s: {**p, "sigma": p["sigma"] + s * hv}, lambda s: S0)
    v = (up - dn) / (2 * hv)
    want_vega, se_vega = v.mean(), v.std(ddof=1) / math.sqrt(v.size)
    t = ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=m, seed=11)
    got_delta = t.delta(call, stock, ob.KouModel(), rate, h_frac=1e-3)
    got_vega = t.vega(call, stock, ob.KouModel(), rate)
    # the device estimate has 1/16 of the variance (16x the paths)
    f = math.sqrt(1 + 1 / 16)
    assert abs(got_delta - want_delta) < 3.5 * se_delta * f, (got_delta, want_delta, se_delta)
    assert abs(got_vega - want_vega) < 3.5 * se_vega * f, (got_vega, want_vega, se_vega)
