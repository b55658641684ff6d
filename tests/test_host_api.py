"""
CPU-only checks of the host layer: the C-ABI library loads and exports every
symbol include/optmc.h declares, and the technique classes keep the reference's
surface (tests modelled on /root/reference/tests/techniques/test_monte_carlo.py,
base/test_greek_mixin.py, base/test_random_utils.py, base/test_iv_mixin.py).
No kernel is launched here.
"""
import math
import os
import re
from unittest.mock import MagicMock, patch

import numpy as np
import pytest

import optpricing_b200 as ob
from optpricing_b200 import _engine
from optpricing_b200.techniques import monte_carlo as mc_mod
from optpricing_b200.techniques.kernels import mc_kernels, path_kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def setup():
    return (ob.Option(strike=100, maturity=1.0, option_type=ob.OptionType.CALL),
            ob.Stock(spot=100), ob.Rate(rate=0.05))


def test_library_exports_every_declared_symbol():
    from optpricing_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "optmc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(optmc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = _engine.load_library()
    bound = {name for name, _, _ in _engine.SIGNATURES}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.optmc_version() >= 100


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors against the C compiler's view of include/optmc.h."""
    import shutil
    import subprocess
    assert ctypes_sizeof(_engine.OptmcModel) == 8 + 16 * 8 + 8 + 8 + 16
    assert ctypes_sizeof(_engine.OptmcSim) == 16 + 24 + 8 + 8 + 8
    assert ctypes_sizeof(_engine.OptmcLevy) == 8 + 8 * _engine.LEVY_NPARAM
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "optmc.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %d %d\\n", sizeof(optmc_model), '
                   'sizeof(optmc_sim), sizeof(optmc_levy), offsetof(optmc_model, lv_table_dev), '
                   'offsetof(optmc_sim, seed), OPTMC_MAX_PEERS, OPTMC_IPC_HANDLE_BYTES);return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert got == [ctypes_sizeof(_engine.OptmcModel), ctypes_sizeof(_engine.OptmcSim),
                   ctypes_sizeof(_engine.OptmcLevy), _engine.OptmcModel.lv_table_dev.offset,
                   _engine.OptmcSim.seed.offset, _engine.MAX_PEERS, _engine.IPC_HANDLE_BYTES]


def ctypes_sizeof(t):
    import ctypes
    return ctypes.sizeof(t)


def test_ctor_follows_reference():
    t = ob.MonteCarloTechnique()
    assert (t.n_paths, t.n_steps, t.antithetic) == (20_000, 100, True)
    assert isinstance(t.rng, np.random.Generator)
    assert ob.MonteCarloTechnique(n_paths=11).n_paths == 12            # monte_carlo.py:58-59
    assert ob.MonteCarloTechnique(n_paths=11, antithetic=False).n_paths == 11
    a = ob.AmericanMonteCarloTechnique()
    assert (a.n_paths, a.n_steps, a.antithetic, a.lsm_degree) == (20_000, 100, True, 2)
    assert ob.AmericanMonteCarloTechnique(n_paths=7).n_paths == 8
    with pytest.raises(TypeError):
        ob.MonteCarloTechnique(20_000)                                   # keyword-only


def test_mc_model_support_check(setup):
    option, stock, rate = setup
    bad = MagicMock()
    bad.supports_sde = False
    bad.is_pure_levy = False
    bad.name = "Unsupported"
    with pytest.raises(TypeError, match="does not support simulation"):
        ob.MonteCarloTechnique().price(option, stock, bad, rate)
    with pytest.raises(TypeError, match="requires an SDE model"):
        ob.AmericanMonteCarloTechnique().price(option, stock, bad, rate)


@patch("optpricing_b200.techniques.monte_carlo.MonteCarloTechnique._simulate_sde_path")
def test_sde_path_dispatcher(mock_sim, setup):
    mock_sim.return_value = np.array([100.0, 120.0])
    option, stock, rate = setup
    res = ob.MonteCarloTechnique().price(option, stock, ob.BSMModel(), rate)
    mock_sim.assert_called_once()
    assert res.price == pytest.approx(10.0 * math.exp(-0.05))
    assert isinstance(res, ob.PricingResult) and isinstance(res.price, float)


@patch("optpricing_b200.techniques.monte_carlo.MonteCarloTechnique._simulate_levy_terminal")
def test_levy_terminal_dispatcher(mock_levy, setup):
    mock_levy.return_value = np.array([100.0])
    option, stock, rate = setup
    levy = MagicMock()
    levy.supports_sde = False
    levy.is_pure_levy = True
    levy.has_exact_sampler = False
    ob.MonteCarloTechnique().price(option, stock, levy, rate)
    mock_levy.assert_called_once()


def test_exact_sampler_dispatcher(setup):
    option, stock, rate = setup
    cev = MagicMock()
    cev.supports_sde = True
    cev.has_exact_sampler = True
    cev.sample_terminal_spot.return_value = np.array([100.0])
    ob.MonteCarloTechnique().price(option, stock, cev, rate)
    cev.sample_terminal_spot.assert_called_once()


def test_levy_terminal_formula():
    """monte_carlo.py:259-280 with a model whose raw cf is exp(t*psi)."""
    m = MagicMock()
    m.name = "fake"
    m.sample_terminal_log_return.return_value = np.array([0.0, 0.1])
    m.raw_cf.return_value = lambda u: np.exp(0.3 + 0j)
    got = ob.MonteCarloTechnique()._simulate_levy_terminal(m, 100.0, 0.05, 0.01, 2.0)
    want = 100.0 * np.exp((0.05 - 0.01) * 2.0 - 0.3 + np.array([0.0, 0.1]))
    np.testing.assert_allclose(got, want, rtol=1e-15)


@pytest.mark.parametrize("model, expected, keys", [
    (ob.BSMModel(), mc_kernels.bsm_kernel, {"r", "q", "dt", "sigma"}),
    (ob.HestonModel(), mc_kernels.heston_kernel,
     {"r", "q", "dt", "kappa", "theta", "rho", "vol_of_vol", "v0"}),
    (ob.MertonJumpModel(), mc_kernels.merton_kernel,
     {"r", "q", "dt", "sigma", "lambda_", "mu_j", "sigma_j"}),
    (ob.BatesModel(), mc_kernels.bates_kernel,
     {"r", "q", "dt", "kappa", "theta", "rho", "vol_of_vol", "v0", "lambda_", "mu_j", "sigma_j"}),
    (ob.KouModel(), mc_kernels.kou_kernel,
     {"r", "q", "dt", "sigma", "lambda_", "p_up", "eta1", "eta2"}),
    (ob.SABRModel(), mc_kernels.sabr_kernel, {"r", "q", "dt", "alpha", "beta", "rho", "v0"}),
    (ob.SABRJumpModel(), mc_kernels.sabr_jump_kernel,
     {"r", "q", "dt", "alpha", "beta", "rho", "v0", "lambda_", "mu_j", "sigma_j"}),
])
def test_get_sde_kernel_selector(model, expected, keys):
    fn, kp = ob.MonteCarloTechnique()._get_sde_kernel_and_params(model=model, r=0.05, q=0.01,
                                                                 dt=0.01)
    assert fn is expected
    assert set(kp) == keys
    assert (kp["r"], kp["q"], kp["dt"]) == (0.05, 0.01, 0.01)


def test_v0_resolution():
    t = ob.MonteCarloTechnique()
    _, kp = t._get_sde_kernel_and_params(ob.HestonModel(), 0.05, 0.0, 0.01, v0=0.09)
    assert kp["v0"] == 0.09
    _, kp = t._get_sde_kernel_and_params(ob.HestonModel(), 0.05, 0.0, 0.01)
    assert kp["v0"] == 0.04
    _, kp = t._get_sde_kernel_and_params(ob.SABRModel(), 0.05, 0.0, 0.01)
    assert kp["v0"] == 0.5                                               # alpha, monte_carlo.py:136
    fn, kp = ob.AmericanMonteCarloTechnique()._get_path_kernel_and_params(
        model=ob.HestonModel(), r=0.05, q=0.01, dt=0.01, v0=0.04)
    assert fn is path_kernels.heston_path_kernel and kp["v0"] == 0.04


def test_shard_draws_is_a_partition():
    for n in (0, 1, 7, 8, 1 << 23, 12345677):
        for w in (1, 2, 3, 8):
            pieces = [mc_mod.shard_draws(n, r, w) for r in range(w)]
            assert pieces[0][0] == 0 and sum(c for _, c in pieces) == n
            for (o1, c1), (o2, _) in zip(pieces, pieces[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in pieces) - min(c for _, c in pieces) <= 1


def test_summarize_and_control_variate():
    rng = np.random.default_rng(3)
    c = rng.lognormal(size=4000) * 100
    p = np.maximum(c - 100, 0)
    m = [p.size, p.sum(), (p * p).sum(), c.sum(), (c * c).sum(), (p * c).sum()]
    s = mc_mod.summarize(m, 0.05, 1.0, 165.0)
    disc = math.exp(-0.05)
    assert s["price"] == pytest.approx(disc * p.mean(), rel=1e-13)
    assert s["stderr"] == pytest.approx(disc * p.std(ddof=1) / math.sqrt(p.size), rel=1e-10)
    beta = np.cov(p, c, ddof=0)[0, 1] / c.var()
    assert s["cv_beta"] == pytest.approx(beta, rel=1e-9)
    assert s["cv_price"] == pytest.approx(disc * (p.mean() - beta * (c.mean() - 165.0)), rel=1e-9)
    assert s["cv_stderr"] < s["stderr"]


def test_summarize_without_samples_is_nan():
    st = mc_mod.summarize([0, 0, 0, 0, 0, 0], 0.05, 1.0, 105.0)
    assert math.isnan(st["price"]) and st["n_samples"] == 0.0
    assert np.all(np.isnan(mc_mod.prices_from_moments(np.zeros((2, 6)), 0.05, 1.0, 105.0, False)))


def test_control_mean():
    f = mc_mod.control_mean
    assert f("heston", {}, 100.0, 0.05, 0.01, 2.0, 10) == pytest.approx(100 * math.exp(0.08))
    assert f("sabr", {}, 100.0, 0.05, 0.01, 1.0, 4) == pytest.approx(100 * 1.01**4)
    p = {"lambda": 0.4, "mu_j": -0.1, "sigma_j": 0.15}
    k = math.exp(-0.1 + 0.5 * 0.15**2) - 1
    want = 100 * ((1 + (0.04 - 0.4 * k) * 0.25) * math.exp(0.4 * k * 0.25)) ** 4
    assert f("sabr_jump", p, 100.0, 0.05, 0.01, 1.0, 4) == pytest.approx(want)


# ---- mixins, with an analytic pricer (reference: tests/techniques/base/) ------
class LinearPricer(ob.BaseTechnique, ob.GreekMixin, ob.IVMixin):
    """price = 2 S - 3 K + 5 sigma + 7 T + 11 r: every greek is known exactly."""

    def __init__(self, with_rng):
        if with_rng:
            self.rng = np.random.default_rng(0)
        self.calls = 0

    def price(self, option, stock, model, rate, **kw):
        self.calls += 1
        noise = self.rng.standard_normal() if hasattr(self, "rng") else 0.0
        v = (2 * stock.spot - 3 * option.strike + 5 * model.params.get("sigma", 0.0)
             + 7 * option.maturity + 11 * rate.get_rate(option.maturity) + noise)
        return ob.PricingResult(price=v)


@pytest.mark.parametrize("with_rng", [True, False])
def test_greek_mixin_central_differences(setup, with_rng):
    option, stock, rate = setup
    t, m = LinearPricer(with_rng), ob.BSMModel()
    assert t.delta(option, stock, m, rate) == pytest.approx(2.0, abs=1e-6)
    assert t.gamma(option, stock, m, rate) == pytest.approx(0.0, abs=1e-3)
    assert t.vega(option, stock, m, rate) == pytest.approx(5.0, abs=1e-6)
    assert t.theta(option, stock, m, rate) == pytest.approx(-7.0, abs=1e-4)
    assert t.rho(option, stock, m, rate) == pytest.approx(11.0, abs=1e-6)
    assert t.calls == 2 + 3 + 2 + 2 + 2
    assert np.isnan(t.vega(option, stock, ob.HestonModel(), rate))       # greek_mixin.py:156-157


def test_crn_restores_generator_state():
    rng = np.random.default_rng(5)
    with ob.crn(rng):
        a = rng.standard_normal(4)
    with ob.crn(rng):
        b = rng.standard_normal(4)
    assert np.array_equal(a, b)
    with pytest.raises(RuntimeError):
        with ob.crn(rng):
            rng.standard_normal()
            raise RuntimeError
    assert np.array_equal(rng.standard_normal(4), a)


class BSMAnalytic(ob.BaseTechnique, ob.IVMixin):
    def price(self, option, stock, model, rate, **kw):
        from scipy.stats import norm
        S, K, T, r = stock.spot, option.strike, option.maturity, rate.get_rate()
        s = model.params["sigma"]
        d1 = (math.log(S / K) + (r + 0.5 * s * s) * T) / (s * math.sqrt(T))
        return ob.PricingResult(price=S * norm.cdf(d1) - K * math.exp(-r * T) * norm.cdf(d1 - s * math.sqrt(T)))


def test_implied_volatility_roundtrip(setup):
    option, stock, rate = setup
    t = BSMAnalytic()
    target = t.price(option, stock, ob.BSMModel({"sigma": 0.27}), rate).price
    assert t.implied_volatility(option, stock, ob.BSMModel(), rate, target) == pytest.approx(0.27, abs=1e-5)
    assert np.isnan(t.implied_volatility(option, stock, ob.BSMModel(), rate, 1e9))
    assert ob.IVMixin._secant_iv(lambda x: x * x - 0.04, 0.3, 1e-10, 100) == pytest.approx(0.2)


def test_atoms_and_models_validate():
    with pytest.raises(ValueError):
        ob.Option(strike=-1, maturity=1.0, option_type=ob.OptionType.CALL)
    with pytest.raises(ValueError):
        ob.Stock(spot=0)
    assert ob.Rate(rate=lambda t: 0.02 + 0.01 * t).get_rate(2.0) == pytest.approx(0.04)
    with pytest.raises(ValueError):
        ob.HestonModel({"kappa": 1.0, "theta": 0.04, "rho": 1.5, "vol_of_vol": 0.3})
    m = ob.BSMModel().with_params(sigma=0.3)
    assert m.params["sigma"] == 0.3 and m == ob.BSMModel({"sigma": 0.3})
    kinds = [ob.model_kind(c()) for c in (ob.BSMModel, ob.HestonModel, ob.MertonJumpModel,
                                          ob.BatesModel, ob.KouModel, ob.SABRModel,
                                          ob.SABRJumpModel)]
    assert kinds == ["bsm", "heston", "merton", "bates", "kou", "sabr", "sabr_jump"]


def test_engine_fails_loudly_without_gpu(setup):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    option, stock, rate = setup
    with pytest.raises(_engine.EngineError, match="no CPU fallback"):
        ob.MonteCarloTechnique(n_paths=100, n_steps=2).price(option, stock, ob.BSMModel(), rate)


def test_philox_reference_known_answers():
    from philox_ref import KAT, philox4x32_10
    for c, k, want in KAT:
        got = philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], *k)[0]
        assert [int(x) for x in got] == list(want)


def test_price_frame_validates_without_a_gpu():
    """price_frame (the MC counterpart of price_options_vectorized): bad option types raise,
    an empty frame prices to an empty array -- both before any device work."""
    import pandas as pd
    t = ob.MonteCarloTechnique(n_paths=100, n_steps=2, seed=1)
    stock, rate = ob.Stock(100.0), ob.Rate(0.05)
    bad = pd.DataFrame({"strike": [100.0], "maturity": [1.0], "optionType": ["straddle"]})
    with pytest.raises(ValueError, match="optionType"):
        t.price_frame(bad, stock, ob.BSMModel(), rate)
    empty = pd.DataFrame({"strike": [], "maturity": [], "optionType": []})
    assert t.price_frame(empty, stock, ob.BSMModel(), rate).shape == (0,)


def test_box_muller_log_table_is_positive_on_the_46_bit_grid():
    """normal_parts_fast (csrc/philox.cuh) takes -2 ln u from a 512-entry table + degree-4
    polynomial with an absolute error of ~5e-16 and feeds it to a square root WITHOUT a floor
    at zero.  That is safe on the pair mapping's grid u = (x + 1/2) 2^-46: an exact emulation
    (neg2_log_poly, same constants, fused multiply-adds in exact rational arithmetic) stays
    positive and accurate for the largest uniforms of the top binade."""
    from fractions import Fraction as F

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    log_n, c_1_12, c_n2ln2 = 512, 1.0 / 12.0, -1.3862943611198906188

    def neg2_log(m, kfac):
        i = int((m - 1.0) * log_n)
        c = 1.0 + (i + 0.5) / log_n
        tx, ty = -2.0 / c, -2.0 * math.log(c)                              # optmc.cu: RngBase.logt
        r = fma(m, tx, 2.0)
        p = fma(r, 0.03125, c_1_12)
        p = fma(p, r, 0.25)
        p = fma(p, r, 1.0)
        return fma(p, r, fma(kfac, c_n2ln2, ty))

    # u = m 2^kfac with kfac = -1 in the top binade; x = 2^46 - j  ->  m = 2 - (2 j - 1) 2^-46
    js = list(range(1, 400)) + [2 ** e for e in range(9, 40)]
    for j in js:
        m = 2.0 - (2 * j - 1) * 2.0 ** -46
        u = 0.5 * m
        got = neg2_log(m, -1.0)
        assert got > 0.0
        assert got == pytest.approx(-2.0 * math.log(u), abs=1e-15)
    # even the two largest uniforms of a 64-bit grid (Kou's jump sizes draw from one) stay
    # positive now that the table's y carries no exponent offset (round 1's y = -2 ln c + 24 ln 2
    # made them come out at -2e-16, the reason that generator carried a floor)
    top = [neg2_log(2.0 - k * 2.0 ** -52, -1.0) for k in (1, 2)]
    assert all(0.0 < v < 1e-15 for v in top), top


def test_default_rng_mode_switches():
    """overlay.install(rng_mode=...) / OPTPRICING_B200_RNG_MODE choose the draw source of
    techniques constructed without rng_mode= (the mode in which the reference's lucky-seed
    acceptance test passes by construction is "numpy")."""
    from optpricing_b200.techniques import monte_carlo as m
    assert ob.MonteCarloTechnique().rng_mode == "philox"
    assert ob.AmericanMonteCarloTechnique().rng_mode == "philox"
    m.set_default_rng_mode("numpy")
    try:
        assert ob.MonteCarloTechnique().rng_mode == "numpy"
        assert ob.AmericanMonteCarloTechnique().rng_mode == "numpy"
        assert ob.MonteCarloTechnique(rng_mode="philox").rng_mode == "philox"
    finally:
        m.set_default_rng_mode(None)
    with patch.dict(os.environ, {"OPTPRICING_B200_RNG_MODE": "numpy"}):
        assert ob.MonteCarloTechnique().rng_mode == "numpy"
    assert ob.MonteCarloTechnique().rng_mode == "philox"
    with pytest.raises(ValueError):
        m.set_default_rng_mode("sobol")


def test_poisoned_moments_raise():
    """A rank that never arrived at k_peer_allreduce leaves NaN 0x7FF8DEADDEADDEAD, not numbers."""
    mom = np.array([8.0, 1.0, 2.0, 3.0, 4.0, 5.0])
    mc_mod.summarize(mom, 0.05, 1.0, 100.0)
    bad = mom.copy()
    bad.view(np.uint64)[1] = 0x7FF8DEADDEADDEAD
    with pytest.raises(_engine.EngineError, match="peer all-reduce"):
        mc_mod.summarize(bad, 0.05, 1.0, 100.0)
    with pytest.raises(_engine.EngineError, match="peer all-reduce"):
        mc_mod.prices_from_moments(np.stack([mom, bad]), 0.05, 1.0, 100.0, False)
    plain_nan = mom.copy()
    plain_nan[1] = float("nan")                       # an ordinary NaN is not poison
    assert math.isnan(mc_mod.summarize(plain_nan, 0.05, 1.0, 100.0)["price"])


def test_price_with_greeks_validates_its_arguments(setup):
    opt, st, rt = setup
    with pytest.raises(ValueError, match="unknown greeks"):
        ob.MonteCarloTechnique().price_with_greeks(opt, st, ob.BSMModel(), rt, greeks=("delta", "speed"))


def test_pricings_are_serialised_by_a_reentrant_lock():
    import threading
    order = []

    @_engine.serialized
    def inner(tag):
        order.append(("in", tag))
        return tag

    @_engine.serialized
    def outer(tag):
        order.append(("enter", tag))
        inner(tag)                                   # re-entrant: greeks call price()
        order.append(("leave", tag))

    ts = [threading.Thread(target=outer, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in range(0, len(order), 3):                # no interleaving between threads
        tags = {t for _, t in order[k:k + 3]}
        assert len(tags) == 1 and [w for w, _ in order[k:k + 3]] == ["enter", "in", "leave"]


def test_bench_sass_constants_match_the_built_kernel():
    """bench.py's roofline multiplies path-steps/s by static SASS counts of the headline
    kernel's step loop: hold them to what tools/sass_stats.py reads off the built library."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    from optpricing_b200.build import build
    build()
    sys.path.insert(0, ROOT)
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_stats.py"),
                          r"k_european_wide<1, true, false, false, 768>"], capture_output=True,
                         text=True, cwd=ROOT, timeout=600).stdout
    loops = re.findall(r"loop \S+: (\d+) instr, fp64-pipe (\d+) .*?MUFU (\d+), .*?reg reads (\d+)", out)
    main = [tuple(map(int, l)) for l in loops if int(l[2]) == 4]    # two pair-steps: four roots
    assert main, out[-2000:]
    instr, fp64, _, reads = min(main)                              # the whole-call loop body
    sass = bench.SASS["heston"]                                    # per PATH-step: body / 4
    assert fp64 / 4 == pytest.approx(sass["i64"], rel=0.01)
    assert instr / 4 == pytest.approx(sass["instr"], rel=0.04)
    assert reads / 4 == pytest.approx(sass["reads"], rel=0.04)


def test_no_cp_async_with_unset_uniform_registers_in_the_built_library():
    """ptxas 12.9 has encoded hinted cp.async copies of one k_lsmc_persistent instantiation as
    `LDGSTS [R+UR0], desc[UR1]` with neither uniform register ever written; the GPU reports "an
    illegal instruction was encountered".  The trigger is a code-generation detail (an unrolled
    request loop), so the built library is scanned for that addressing form."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None or not os.path.exists(_engine.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run(["cuobjdump", "-sass", _engine.LIB_PATH], capture_output=True, text=True,
                          check=True).stdout
    copies = [ln for ln in sass.splitlines() if "LDGSTS" in ln]
    assert len(copies) >= 30                      # the ring kernels are in the library
    bad = [ln.strip() for ln in copies if re.search(r"\[R\d+\+UR\d+", ln)]
    assert not bad, bad[:3]


def test_theta_from_one_simulation_guards():
    """GBM theta is read off one simulation only where the identity behind it holds: the
    maturity bumps keep the step count (steps_per_year may round T +- h to another grid), the
    device RNG is in use, and the model is GBM; everything else goes through the mixin's two
    re-pricings (here: a model with its own exact sampler, counted)."""
    t = ob.MonteCarloTechnique(n_paths=100, n_steps=10)
    assert t._theta_steps_ok(1.0, 1e-5)
    t = ob.MonteCarloTechnique(n_paths=100, n_steps=10, steps_per_year=252)
    assert t._theta_steps_ok(1.0, 1e-5)
    assert not t._theta_steps_ok(10.5 / 252 - 1e-6, 1e-5)       # T + h rounds to 11 steps, T to 10
    model = MagicMock()
    model.supports_sde = True
    model.has_exact_sampler = True
    model.name = "mock"
    model.sample_terminal_spot.side_effect = lambda S0, r, T, n: np.full(n, S0 * math.exp(r * T) * 1.01)
    t = ob.MonteCarloTechnique(n_paths=1000, n_steps=4, seed=3)
    opt = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
    th = t.theta(opt, ob.Stock(spot=100.0), model, ob.Rate(rate=0.05))
    assert model.sample_terminal_spot.call_count == 2              # T + h and T - h
    assert sorted(c.args[2] for c in model.sample_terminal_spot.call_args_list) == pytest.approx(
        [1.0 - 1e-5, 1.0 + 1e-5])
    # deterministic terminal S0 e^{rT} 1.01: price = 1.01 S0 - K e^{-rT}, theta = -r K e^{-rT}
    assert th == pytest.approx(-0.05 * 100.0 * math.exp(-0.05), rel=1e-5)


def test_pow_beta_table_scheme_is_accurate_by_emulation():
    """pow_beta (csrc/models.cuh; the SABR scenario kernel): s^beta = 2^(e beta) c_i^beta (1 + rho)^beta
    with the log table's nodes c_i, the binomial series to degree 4 in r' = -2 rho with the host's
    coefficients (optmc.cu) and fused multiply-adds.  Emulated with exact rational FMAs over spots
    from 1e-8 (the floor) to 1e12 and several beta: relative error below 3e-15 (a few ulp: three roundings
    of the product, one of each table entry) -- the scenario
    prices then equal bumped re-pricings far inside the 1e-9 the GPU test asserts."""
    from fractions import Fraction as F

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    rng = np.random.default_rng(8)
    spots = np.concatenate([10.0 ** rng.uniform(-8, 12, 300), [1e-8, 1.0, 100.0, 2.0 - 2.0 ** -52,
                                                              1.0 + 2.0 ** -52, 65536.0]])
    for beta in (0.8, 0.5, 0.3, 0.999, 0.05):
        pw = [-beta / 2.0, beta * (beta - 1.0) / 8.0, -beta * (beta - 1.0) * (beta - 2.0) / 48.0,
              beta * (beta - 1.0) * (beta - 2.0) * (beta - 3.0) / 384.0]
        worst = 0.0
        for s in spots:
            m, e = math.frexp(float(s))            # s = m 2^e, m in [0.5, 1)
            m, e = 2.0 * m, e - 1                  # m in [1, 2): the mantissa the kernel rebuilds
            i = int((m - 1.0) * 512)
            c = 1.0 + (i + 0.5) / 512
            tx = -2.0 / c                          # optmc.cu: RngBase.logt[i].x
            r = fma(m, tx, 2.0)
            p = fma(r, pw[3], pw[2])
            p = fma(p, r, pw[1])
            p = fma(p, r, pw[0])
            p = fma(p, r, 1.0)
            tc = (-2.0 / tx) ** beta               # european.cuh: pow(-2 / logt[i].x, beta)
            te = 2.0 ** (beta * e)                 # exp2(beta (k - 64)), k = e + 64 in 0 .. 127
            assert 0 <= e + 64 <= 127
            got = (te * tc) * p
            worst = max(worst, abs(got / float(s) ** beta - 1.0))
        assert worst < 3e-15, (beta, worst)
