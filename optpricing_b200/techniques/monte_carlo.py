"""
MonteCarloTechnique on the B200 engine: a drop-in for
/root/reference/src/optpricing/techniques/monte_carlo.py.

Public surface, defaults and error messages follow the reference
(monte_carlo.py:36-116); what changed is underneath:

* ``rng_mode="philox"`` (default).  ``price`` makes ONE call into liboptmc.so:
  Philox normals / Poisson / jump sizes are generated in registers, the payoff
  moments come back as 48 bytes.  No increments are materialised, so 2^24 x 252
  runs in milliseconds where the reference cannot even allocate its arrays.  The
  Philox key is one draw from ``self.rng``, hence ``seed=`` reproducibility,
  stream advance per ``price`` call and common random numbers under ``crn`` all
  behave as in the reference (greek_mixin.py:63-64 keys CRN on ``self.rng``).
* ``rng_mode="numpy"``.  The reference's own draw sequence
  (monte_carlo.py:222-240) on the host, fed to the GPU fed-draw kernels: prices
  equal the reference's bit for bit up to 1e-12 for the same ``seed`` (and the
  same ``np.random.seed`` for jump sizes).  Host-RNG bound; for parity and audits.
* ``correlation="reference"`` reproduces the reference's European "double
  correlation" of the two Brownian drivers (monte_carlo.py:230-231 followed by
  mc_kernels.py:67); ``"independent"`` gives corr(dW_S, dW_v) = rho.
* ``precision="fp32"`` (optional, device RNG only): the step loop in single
  precision on the fp32 pipe, the per-path finish and the moments in fp64.
* ``report_stats=True`` puts the standard error, a 95 % interval and the sample
  count of the estimate into ``PricingResult.greeks`` (``last_stats`` always has them).
* ``control_variate=True`` adds the discounted-terminal-spot control, whose mean
  under the simulated discrete scheme is known in closed form.
* Under an initialised ``torch.distributed`` group the draws are partitioned by
  rank (disjoint Philox counters) and the moments are all-reduced once.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Any

import numpy as np

from optpricing_b200 import _engine
from optpricing_b200.atoms import Option, OptionType
from optpricing_b200.atoms import is_call as _is_call
from optpricing_b200.models import levy_kind, model_kind

from .base import BaseTechnique, GreekMixin, IVMixin, PricingResult, crn
from .kernels import mc_kernels

_KERNELS = {
    "bsm": mc_kernels.bsm_kernel, "heston": mc_kernels.heston_kernel,
    "merton": mc_kernels.merton_kernel, "bates": mc_kernels.bates_kernel,
    "sabr": mc_kernels.sabr_kernel, "sabr_jump": mc_kernels.sabr_jump_kernel,
    "kou": mc_kernels.kou_kernel, "dupire": mc_kernels.dupire_kernel,
}


def shard_draws(n_draws: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous partition of draw ids [0, n_draws) -> (offset, count) of ``rank``."""
    lo = n_draws * rank // world
    hi = n_draws * (rank + 1) // world
    return lo, hi - lo


def control_mean(kind: str, params: dict, S0: float, r: float, q: float, T: float,
                 n_steps: int) -> float:
    """
    E[S_T] under the *discrete* scheme the kernels simulate.  The log-Euler
    models are exact one-step martingales (jump compensators match E[e^J]), so
    E[S_T] = S0 e^{(r-q)T}; SABR is Euler on S itself (mc_kernels.py:211,257).
    """
    if kind in ("cev", "lognormal"):
        return S0 * math.exp(r * T)               # cev.py:49-96 takes no dividend yield
    if kind in ("nig", "hyperbolic", "cgmy"):
        # the reference's samplers of these laws are not the time-T marginals of the
        # process raw_cf describes (nig.py:163-166 draws IG with shape 1), so the
        # martingale-corrected mean is not S0 e^{(r-q)T}: no control variate
        return float("nan")
    dt = T / n_steps
    if kind == "sabr":
        return S0 * (1.0 + (r - q) * dt) ** n_steps
    if kind == "sabr_jump":
        k = math.exp(params["mu_j"] + 0.5 * params["sigma_j"] ** 2) - 1.0
        lam = params["lambda"]
        return S0 * ((1.0 + (r - q - lam * k) * dt) * math.exp(lam * k * dt)) ** n_steps
    return S0 * math.exp((r - q) * T)


def _group_all_reduce(dist):
    """The all-reduce of a partitioned pricing: the engine's own kernel over NVLink peer memory
    on an NCCL group (csrc/lsmc.cuh: k_peer_allreduce), torch.distributed's otherwise."""
    try:
        pick = getattr(_engine.get_engine(), "group_all_reduce", None)
    except _engine.EngineError:          # no GPU in this process (gloo tests of the host logic)
        pick = None
    return pick() if pick is not None else dist.all_reduce


_POISON = np.uint64(0x7FF8DEADDEADDEAD)


def _check_reduced(mom: np.ndarray) -> None:
    """A rank that never arrived at a peer all-reduce leaves poison, not numbers."""
    m = np.ascontiguousarray(mom, dtype=np.float64)
    if np.isnan(m).any() and (m.view(np.uint64) == _POISON).any():
        raise _engine.EngineError("peer all-reduce timed out: a rank of the group did not make "
                                  "this call (the techniques must be called by all ranks, SPMD)")


def summarize(moments, r: float, T: float, c_mean: float) -> dict:
    """Price, standard error and control-variate estimate from (n, Sp, Spp, Sc, Scc, Spc)."""
    n, sp, spp, sc, scc, spc = (float(x) for x in moments)
    if sp != sp or sc != sc:
        _check_reduced(np.asarray(moments))
    disc = math.exp(-r * T)
    if n <= 0.0:        # no samples: the mean of an empty array, as np.mean gives the reference
        return {"n_samples": 0.0, "price": float("nan"), "stderr": float("nan"),
                "control_mean": float("nan"), "control_expected": c_mean,
                "control_stderr": float("nan")}
    mp, mc = sp / n, sc / n
    var_p = max(spp / n - mp * mp, 0.0) * (n / (n - 1.0) if n > 1 else 1.0)
    out = {"n_samples": n, "price": disc * mp, "stderr": disc * math.sqrt(var_p / n),
           "control_mean": mc, "control_expected": c_mean}
    var_c = scc / n - mc * mc
    cov = spc / n - mp * mc
    out["control_stderr"] = math.sqrt(max(var_c, 0.0) / max(n - 1.0, 1.0))
    if var_c > 0.0:
        beta = cov / var_c
        resid = max(var_p - beta * cov * (n / (n - 1.0) if n > 1 else 1.0), 0.0)
        out.update(cv_beta=beta, cv_price=disc * (mp - beta * (mc - c_mean)),
                   cv_stderr=disc * math.sqrt(resid / n))
    return out


def prices_from_moments(mom: np.ndarray, r: float, T: float, c_mean: float, use_cv: bool):
    """Vectorised ``summarize`` for the price only: mom is (n, 6)."""
    _check_reduced(mom)
    n, sp, spp, sc, scc, spc = (mom[:, k] for k in range(6))
    disc = math.exp(-r * T)
    with np.errstate(divide="ignore", invalid="ignore"):      # n = 0 rows price to nan
        mp, mc = sp / n, sc / n
    price = disc * mp
    if not (use_cv and math.isfinite(c_mean)):
        return price
    var_c = scc / n - mc * mc
    cov = spc / n - mp * mc
    ok = var_c > 0.0
    beta = np.where(ok, cov / np.where(ok, var_c, 1.0), 0.0)
    return np.where(ok, disc * (mp - beta * (mc - c_mean)), price)


class LazyStats:
    """``last_stats_many``: the per-contract statistics dict is built on access, so a
    sweep over hundreds of contracts does not pay for dictionaries nobody reads."""

    def __init__(self, n):
        self._rows = [None] * n

    def _set(self, i, row, r, T, c_mean, seed):
        self._rows[i] = (row, r, T, c_mean, seed)

    def __len__(self):
        return len(self._rows)

    def __getitem__(self, i):
        row, r, T, c_mean, seed = self._rows[i]
        st = summarize(row, r, T, c_mean)
        st["seed"] = seed
        return st

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class DeviceTerminal:
    """What the Philox path hands from the simulator to ``price``: the payoff
    moments of the hinted contract instead of a host array of terminal spots."""

    def __init__(self, moments, kind, seed):
        self.moments, self.kind, self.seed = moments, kind, seed


def default_rng_mode() -> str:
    """Draw source of a technique constructed without ``rng_mode=``: "philox" (device RNG)
    unless ``overlay.install(rng_mode=...)`` or OPTPRICING_B200_RNG_MODE says "numpy" -- the
    mode in which a seeded technique reproduces the reference's own numbers (and its
    lucky-seed acceptance test, tests/techniques/test_monte_carlo.py:111-126) to 1e-12."""
    return _DEFAULT_RNG_MODE[0] or os.environ.get("OPTPRICING_B200_RNG_MODE", "philox")


_DEFAULT_RNG_MODE = [None]


def set_default_rng_mode(mode: str | None) -> None:
    if mode not in (None, "philox", "numpy"):
        raise ValueError("rng_mode must be 'philox' or 'numpy'")
    _DEFAULT_RNG_MODE[0] = mode


class MonteCarloTechnique(BaseTechnique, GreekMixin, IVMixin):
    def __init__(self, *, n_paths: int = 20_000, n_steps: int = 100, antithetic: bool = True,
                 seed: int | None = None, rng_mode: str | None = None,
                 correlation: str = "reference", control_variate: bool = False,
                 distributed: bool | None = None, precision: str = "fp64",
                 steps_per_year: int | None = None, report_stats: bool = False):
        if antithetic and n_paths % 2 != 0:
            n_paths += 1                       # monte_carlo.py:58-59
        rng_mode = rng_mode or default_rng_mode()
        if rng_mode not in ("philox", "numpy"):
            raise ValueError("rng_mode must be 'philox' or 'numpy'")
        if correlation not in ("reference", "independent"):
            raise ValueError("correlation must be 'reference' or 'independent'")
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        if precision == "fp32" and rng_mode != "philox":
            raise ValueError("the fp32 mode exists for the device RNG only (rng_mode='philox')")
        self.n_paths = n_paths
        self.n_steps = n_steps
        self.antithetic = antithetic
        self.rng = np.random.default_rng(seed)
        self.rng_mode = rng_mode
        self.correlation = correlation
        self.control_variate = control_variate
        self.distributed = distributed
        self.precision = precision
        self.steps_per_year = steps_per_year
        self.report_stats = report_stats
        self.last_stats: dict = {}
        self._hint = None

    @contextlib.contextmanager
    def _steps_for(self, T: float):
        """``steps_per_year`` (optional): the time grid follows the maturity,
        n_steps = round(steps_per_year * T) -- BASELINE config 3's "n_steps = 252 T" --
        so one technique prices a strike x maturity grid in ONE ``price_many`` call.
        Without it ``n_steps`` is the reference's fixed count."""
        if self.steps_per_year is None:
            yield self.n_steps
            return
        saved = self.n_steps
        self.n_steps = max(1, int(round(self.steps_per_year * T)))
        try:
            yield self.n_steps
        finally:
            self.n_steps = saved

    # ------------------------------------------------------------------ price
    @_engine.serialized
    def price(self, option, stock, model, rate, **kwargs: Any):
        """``technique.price(option, stock, model, rate) -> PricingResult`` (monte_carlo.py:65-116)."""
        if self.steps_per_year is not None and not isinstance(option, (list, tuple, np.ndarray)):
            with self._steps_for(option.maturity):
                return self._price_one(option, stock, model, rate, **kwargs)
        return self._price_one(option, stock, model, rate, **kwargs)

    def _price_one(self, option, stock, model, rate, **kwargs: Any):
        """
        Price one option -> PricingResult (monte_carlo.py:65-116).  A list / tuple /
        array of options (the ``Option | np.ndarray`` the reference's ABC hints at,
        base_technique.py:25-32) is priced by ``price_many`` -> np.ndarray.
        """
        if isinstance(option, (list, tuple, np.ndarray)):
            return self.price_many(option, stock, model, rate, **kwargs)
        if not (model.supports_sde or getattr(model, "is_pure_levy", False)):
            raise TypeError(f"Model '{model.name}' does not support simulation.")
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)

        if getattr(model, "has_exact_sampler", False):      # CEV (:103-104)
            sampler = model.sample_terminal_spot
            if (self._device_single_step(model)
                    and getattr(sampler, "__optmc_device__", False) is True):
                # the engine's own sampler: same draws, but only the moments leave the GPU
                ST = self._device_terminal(model, S0, r, q, T, hint=(float(K), call))
            else:                                           # a model bringing its own sampler
                ST = sampler(S0, r, T, self.n_paths)
        elif getattr(model, "is_pure_levy", False):         # VG/NIG/CGMY/Hyperbolic (:105-106)
            self._hint = (float(K), call)
            try:
                ST = self._simulate_levy_terminal(model, S0, r, q, T)
            finally:
                self._hint = None
        else:
            self._hint = (float(K), call)
            try:
                ST = self._simulate_sde_path(model, S0, r, q, T, **kwargs)
            finally:
                self._hint = None

        if isinstance(ST, DeviceTerminal):
            c_mean = control_mean(ST.kind, model.params, S0, r, q, T, self.n_steps)
            stats = summarize(ST.moments, r, T, c_mean)
            stats["seed"] = ST.seed
            self.last_stats = stats
            use_cv = self.control_variate and "cv_price" in stats and math.isfinite(c_mean)
            price = float(stats["cv_price"] if use_cv else stats["price"])
            if self.report_stats:       # the Monte Carlo error travels with the price
                se = float(stats["cv_stderr"] if use_cv else stats["stderr"])
                return PricingResult(price=price, greeks={
                    "stderr": se, "ci95_low": price - 1.959963984540054 * se,
                    "ci95_high": price + 1.959963984540054 * se, "n_samples": stats["n_samples"]})
            return PricingResult(price=price)

        payoff = np.maximum(ST - K, 0) if call else np.maximum(K - ST, 0)
        self.last_stats = {"n_samples": float(np.size(payoff))}
        return PricingResult(price=float(np.mean(payoff) * math.exp(-r * T)))

    # --------------------------------------------------------------- dispatch
    def _get_sde_kernel_and_params(self, model, r: float, q: float, dt: float, **kwargs: Any):
        """(kernel function, keyword arguments) exactly as monte_carlo.py:118-197 pairs them."""
        p = model.params
        kind = model_kind(model)
        kp: dict[str, Any] = {"r": r, "q": q, "dt": dt}
        if kind in ("sabr", "sabr_jump"):
            kp.update(alpha=p["alpha"], beta=p["beta"], rho=p["rho"],
                      v0=kwargs.get("v0", p.get("alpha")))
        elif kind in ("heston", "bates"):
            kp.update(kappa=p["kappa"], theta=p["theta"], rho=p["rho"],
                      vol_of_vol=p["vol_of_vol"], v0=kwargs.get("v0", p.get("v0")))
        elif kind == "kou":
            kp.update(sigma=p["sigma"], lambda_=p["lambda"], p_up=p["p_up"], eta1=p["eta1"],
                      eta2=p["eta2"])
        elif kind == "merton":
            kp.update(sigma=p["sigma"])
        elif kind == "bsm":
            kp["sigma"] = p.get("sigma")
        # (dupire: the kernel's vol_surface_func is added by the simulator, as in the
        #  reference -- monte_carlo.py:146-147 returns the kernel, :243-244 adds the callable)
        if kind in ("sabr_jump", "bates", "merton"):
            kp.update(lambda_=p["lambda"], mu_j=p["mu_j"], sigma_j=p["sigma_j"])
        return _KERNELS[kind], kp

    def _group(self):
        """(rank, world, all_reduce or None) for the draw partition."""
        if self.distributed is False:
            return 0, 1, None
        try:
            import torch.distributed as dist
        except ImportError:  # pragma: no cover
            return 0, 1, None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0, 1, None
        return dist.get_rank(), dist.get_world_size(), _group_all_reduce(dist)

    # ------------------------------------------------------------- simulators
    def _simulate_sde_path(self, model, S0: float, r: float, q: float, T: float, **kwargs: Any):
        """
        Terminal spots of ``n_paths`` simulated paths (monte_carlo.py:199-257).
        Called directly it returns the host array, originals then antithetic
        partners; called from ``price`` in Philox mode it returns a
        DeviceTerminal carrying the payoff moments, and nothing leaves the GPU.
        """
        if self.rng_mode == "numpy":
            return self._simulate_sde_path_numpy(model, S0, r, q, T, **kwargs)
        T / self.n_steps        # n_steps = 0 raises ZeroDivisionError, as the reference's dt (:209)
        kind = model_kind(model)
        eng = _engine.get_engine()
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        mdl, _keep = self._device_model(eng, kind, model, S0, r, q, T, **kwargs)
        corr = (_engine.CORR_REFERENCE_EUROPEAN if self.correlation == "reference"
                else _engine.CORR_INDEPENDENT_DRAWS)
        rank, world, all_reduce = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed, corr,
                           off, cnt, precision=1 if self.precision == "fp32" else 0)
        if self._hint is None:
            term = eng.buffer("terminal", cnt * (2 if self.antithetic else 1))
            eng.european(mdl, sim, [float(S0)], [1], terminal=term)
            return term[: cnt * (2 if self.antithetic else 1)].cpu().numpy()
        K, call = self._hint
        if all_reduce is None:
            mom = eng.european(mdl, sim, [K], [1 if call else 0])[0]
        elif all_reduce == getattr(eng, "peer_all_reduce", None):
            # simulation, all-reduce over NVLink peer memory and read: one call, one sync
            mom = eng.european_allreduce(mdl, sim, [K], [1 if call else 0])[0]
        else:
            dev = eng.buffer("moments", _engine.NMOM)
            eng.european_async(mdl, sim, [K], [1 if call else 0], dev)
            all_reduce(dev[: _engine.NMOM])
            mom = eng.read(dev, _engine.NMOM)
        return DeviceTerminal(mom, kind, seed)

    def _device_model(self, eng, kind, model, S0, r, q, T, **kwargs):
        """struct optmc_model for one simulation (+ the object keeping its device table alive)."""
        if kind != "dupire":
            return _engine.make_model(kind, model.params, r, q, v0=kwargs.get("v0")), None
        lv = eng.local_vol_table(model.params["vol_surface"], self.n_steps, T / self.n_steps,
                                 math.log(S0), r, q, n_s=kwargs.get("lv_grid", 1024),
                                 half_width=kwargs.get("lv_half_width"))
        return _engine.make_model(kind, model.params, r, q, local_vol=lv), lv

    def _simulate_sde_path_numpy(self, model, S0, r, q, T, **kwargs):
        """The reference's host draw sequence (monte_carlo.py:209-257) on GPU fed kernels."""
        dt = T / self.n_steps
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        kernel, kp = self._get_sde_kernel_and_params(model, r, q, dt, **kwargs)
        sabr = getattr(model, "is_sabr", False)
        args = {"n_paths": num_draws, "n_steps": self.n_steps, **kp}
        args["s0" if sabr else "log_s0"] = S0 if sabr else math.log(S0)
        shape, sd = (num_draws, self.n_steps), math.sqrt(dt)
        if model.has_variance_process:
            dw_v = self.rng.standard_normal(size=shape) * sd
            dw_u = self.rng.standard_normal(size=shape) * sd
            rho = model.params.get("rho", 0)
            args["dw1"] = rho * dw_v + math.sqrt(1 - rho**2) * dw_u
            args["dw2"] = dw_v
        else:
            args["dw"] = self.rng.standard_normal(size=shape) * sd
        if model.has_jumps:
            args["jump_counts"] = self.rng.poisson(lam=model.params["lambda"] * dt, size=shape)
        local_vol = getattr(model, "is_local_vol", False)
        if local_vol:
            args["vol_surface_func"] = model.params["vol_surface"]             # :243-244
        ST = kernel(**args)
        # the reference skips the antithetic pass for local-vol models (:247): n_paths / 2
        # plain samples -- reproduced here so that same-seed prices and sample counts agree
        if self.antithetic and not local_vol:
            for key in ("dw", "dw1", "dw2"):
                if key in args:
                    args[key] = -args[key]
            ST = np.concatenate([ST, kernel(**args)])
        return ST if sabr else np.exp(ST)

    # ------------------------------------------------- batched contracts
    def _philox_sde(self, model) -> bool:
        return (self.rng_mode == "philox" and model.supports_sde
                and not getattr(model, "has_exact_sampler", False)
                and not getattr(model, "is_pure_levy", False))

    def _sweep_async(self, model, S0, r, q, T, strikes, calls, out_dev, seed=None, _aux=None,
                     **kwargs):
        """
        Enqueue ONE simulation to maturity T and the payoff moments of every
        (strike, type) into ``out_dev`` (device, 6 doubles per contract): the
        terminal spots stay in HBM (8 B per path) and a sweep kernel reduces them
        per contract.  Nothing is synchronised or all-reduced here.  Returns
        (kind, seed, keepalive) -- hold ``keepalive`` until the stream has run.
        """
        eng = _engine.get_engine()
        kind = model_kind(model)
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        if seed is None:
            seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        mdl, keep = self._device_model(eng, kind, model, S0, r, q, T, **kwargs)
        corr = (_engine.CORR_REFERENCE_EUROPEAN if self.correlation == "reference"
                else _engine.CORR_INDEPENDENT_DRAWS)
        rank, world, _ = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed, corr,
                           off, cnt, precision=1 if self.precision == "fp32" else 0)
        strikes = np.ascontiguousarray(strikes, dtype=np.float64)
        calls = np.ascontiguousarray(calls, dtype=np.int32)
        term = eng.buffer("terminal", cnt * (2 if self.antithetic else 1))
        for lo in range(0, strikes.size, 64):
            ks, cs = strikes[lo:lo + 64], calls[lo:lo + 64]
            dst = out_dev[lo * _engine.NMOM:(lo + ks.size) * _engine.NMOM]
            if lo == 0:
                eng.european_async(mdl, sim, ks, cs, dst, terminal=term, aux=_aux)
            else:                                   # further chunks reuse the stored terminals
                eng.payoff_sweep(term, cnt, self.antithetic, ks, cs, out=dst)
        return kind, seed, keep

    def _sweep(self, model, S0, r, q, T, strikes, calls, seed=None, _aux=None, **kwargs):
        """One simulation, all contracts of one maturity -> (moments[n, 6], kind, seed)."""
        eng = _engine.get_engine()
        n = len(strikes)
        dev = eng.buffer("moments_many", n * _engine.NMOM)
        kind, seed, _keep = self._sweep_async(model, S0, r, q, T, strikes, calls,
                                              dev[: n * _engine.NMOM], seed=seed, _aux=_aux,
                                              **kwargs)
        _, _, all_reduce = self._group()
        if all_reduce is not None:
            all_reduce(dev[: n * _engine.NMOM])
        return eng.read(dev, n * _engine.NMOM).reshape(n, _engine.NMOM), kind, seed

    @_engine.serialized
    def price_many(self, options, stock, model, rate, **kwargs: Any) -> np.ndarray:
        """
        Prices of many European contracts on one underlying.  Contracts sharing a
        maturity share one simulation (BASELINE config 3: 64 strikes x 4
        maturities = 4 simulations instead of 256).  Equivalent to calling
        ``price`` per contract with common random numbers within a maturity.
        All simulations and sweeps are enqueued back to back; there is ONE
        all-reduce (several ranks) and ONE device-to-host copy for the whole batch.
        """
        options = list(options)
        if not self._philox_sde(model):
            return np.array([self.price(o, stock, model, rate, **kwargs).price for o in options])
        eng = _engine.get_engine()
        prices = np.empty(len(options))
        self.last_stats_many = LazyStats(len(options))
        by_T: dict[float, list[int]] = {}
        for i, o in enumerate(options):
            by_T.setdefault(float(o.maturity), []).append(i)
        S0, q = stock.spot, stock.dividend
        total = len(options)
        dev = eng.buffer("moments_many", total * _engine.NMOM)
        jobs, keep, pos = [], [], 0
        for T, idx in by_T.items():
            r = rate.get_rate(T)
            out = dev[pos * _engine.NMOM:(pos + len(idx)) * _engine.NMOM]
            with self._steps_for(T) as steps:
                kind, seed, k = self._sweep_async(
                    model, S0, r, q, T, [options[i].strike for i in idx],
                    [1 if _is_call(options[i]) else 0 for i in idx], out, **kwargs)
            jobs.append((T, r, idx, pos, kind, seed, steps))
            keep.append(k)
            pos += len(idx)
        _, _, all_reduce = self._group()
        if all_reduce is not None:
            all_reduce(dev[: total * _engine.NMOM])
        mom_all = eng.read(dev, total * _engine.NMOM).reshape(total, _engine.NMOM)
        for T, r, idx, pos, kind, seed, steps in jobs:
            c_mean = control_mean(kind, model.params, S0, r, q, T, steps)
            block = mom_all[pos:pos + len(idx)]
            prices[idx] = prices_from_moments(block, r, T, c_mean, self.control_variate)
            for row, i in zip(block, idx):
                self.last_stats_many._set(i, row, r, T, c_mean, seed)
        return prices

    @_engine.serialized
    def price_frame(self, options_df, stock, model, rate, **kwargs: Any) -> np.ndarray:
        """
        Prices of the contracts of a DataFrame with ``strike``, ``maturity`` and
        ``optionType`` ("call" / "put") columns, aligned with its row order -- the
        Monte Carlo counterpart of ``price_options_vectorized``
        (calibration/vectorized_pricer.py:20-28), which ``VolatilitySurface.
        calculate_model_iv`` (calibration/iv_surface.py:104-150) and the calibrator
        call and which needs a characteristic function; with this, models that have
        none (SABR, SABR-jump) get a model surface.  One simulation per distinct
        maturity (``price_many``).
        """
        strikes = np.asarray(options_df["strike"], dtype=np.float64)
        mats = np.asarray(options_df["maturity"], dtype=np.float64)
        kinds = np.asarray(options_df["optionType"])
        bad = [k for k in np.unique(kinds) if str(k).lower() not in ("call", "put")]
        if bad:
            raise ValueError(f"optionType must be 'call' or 'put', got {bad}")
        options = [Option(strike=float(k), maturity=float(t),
                          option_type=OptionType.CALL if str(c).lower() == "call" else OptionType.PUT)
                   for k, t, c in zip(strikes, mats, kinds)]
        if not options:
            return np.empty(0)
        return self.price_many(options, stock, model, rate, **kwargs)

    # ------------------------------------------------------ fused greeks
    _LOG_KINDS = ("bsm", "heston", "merton", "bates", "kou")

    def _scaling_greeks_ok(self, model) -> bool:
        """Spot / rate bumps can be read off ONE simulation: the log-Euler models'
        log S_T is (log S_0 + r T) plus a term that depends on neither."""
        return (self._philox_sde(model) and model_kind(model) in self._LOG_KINDS
                and self.precision == "fp64")

    def _bumped_prices(self, option, stock, model, rate, spot_factors=(), rate_shifts=(), **kw):
        """
        Prices under S0 -> f S0 and r -> r + h on common random numbers, from one
        simulation: S_T(f S0, r + h) = f e^{hT} S_T(S0, r) path by path, so
          P(f, h) = f e^{-rT} E[(S_T - K')^+]  with  K' = K / (f e^{hT})
        (puts alike).  This is what GreekMixin's re-pricings compute
        (greek_mixin.py:59-73,105-122,249-265), without re-simulating.
        """
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)
        scen = [(f, 0.0) for f in spot_factors] + [(1.0, h) for h in rate_shifts]
        strikes = [K / (f * math.exp(h * T)) for f, h in scen]
        with crn(self.rng):                       # greeks do not advance the stream
            mom, _, _ = self._sweep(model, S0, r, q, T, strikes, [1 if call else 0] * len(scen),
                                    **kw)
        disc = math.exp(-r * T)
        if not self.control_variate:
            return [f * disc * m[1] / m[0] for (f, h), m in zip(scen, mom)]
        # control variate under the scaling: payoff and control both scale by g = f e^{hT},
        # so beta is unchanged and the estimate of the base strike-K' problem scales by f
        # (e^{-(r+h)T} g = f e^{-rT})
        c_mean = control_mean(model_kind(model), model.params, S0, r, q, T, self.n_steps)
        out = []
        for (f, h), m in zip(scen, mom):
            st = summarize(m, r, T, c_mean)
            out.append(f * st.get("cv_price", st["price"]))
        return out

    def _scenario_greeks_ok(self, model) -> bool:
        """SABR kinds: the vol path does not depend on the spot, so bumped-spot paths ride
        through ONE simulation side by side (csrc/european.cuh: k_european_scen)."""
        return (self._philox_sde(model) and model_kind(model) in ("sabr", "sabr_jump")
                and self.precision == "fp64")

    def _scenario_prices(self, option, stock, model, rate, spot_factors, **kw):
        """Prices under S0 -> f S0 for 2 or 3 factors on common random numbers from one
        simulation of a SABR kind: what GreekMixin's re-pricings compute (greek_mixin.py:59-122)."""
        eng = _engine.get_engine()
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        kind = model_kind(model)
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        with crn(self.rng), self._steps_for(T):           # greeks do not advance the stream
            seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
            mdl = _engine.make_model(kind, model.params, r, q, v0=kw.get("v0"))
            corr = (_engine.CORR_REFERENCE_EUROPEAN if self.correlation == "reference"
                    else _engine.CORR_INDEPENDENT_DRAWS)
            rank, world, all_reduce = self._group()
            off, cnt = shard_draws(num_draws, rank, world)
            sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed, corr,
                               off, cnt)
            n = len(spot_factors)
            dev = eng.buffer("moments_scen", 3 * _engine.NMOM)
            eng.european_scen_async(mdl, sim, spot_factors, K, _is_call(option), dev)
            if all_reduce is not None:
                all_reduce(dev[: n * _engine.NMOM])
            mom = eng.read(dev, n * _engine.NMOM).reshape(n, _engine.NMOM)
            out = []
            for f, row in zip(spot_factors, mom):
                c_mean = control_mean(kind, model.params, S0 * f, r, q, T, self.n_steps)
                st = summarize(row, r, T, c_mean)
                use_cv = self.control_variate and "cv_price" in st and math.isfinite(c_mean)
                out.append(float(st["cv_price"] if use_cv else st["price"]))
        return out

    @_engine.serialized
    def delta(self, option, stock, model, rate, h_frac: float = 1e-3, **kwargs: Any) -> float:
        if self._scenario_greeks_ok(model):
            up, dn = self._scenario_prices(option, stock, model, rate, (1 + h_frac, 1 - h_frac),
                                           **kwargs)
            return (up - dn) / (2 * stock.spot * h_frac)
        if not self._scaling_greeks_ok(model):
            return super().delta(option, stock, model, rate, h_frac, **kwargs)
        up, dn = self._bumped_prices(option, stock, model, rate, (1 + h_frac, 1 - h_frac), **kwargs)
        return (up - dn) / (2 * stock.spot * h_frac)

    @_engine.serialized
    def gamma(self, option, stock, model, rate, h_frac: float = 1e-3, **kw: Any) -> float:
        if self._scenario_greeks_ok(model):
            up, mid, dn = self._scenario_prices(option, stock, model, rate,
                                                (1 + h_frac, 1.0, 1 - h_frac), **kw)
            h = stock.spot * h_frac
            return (up - 2 * mid + dn) / (h * h)
        if not self._scaling_greeks_ok(model):
            return super().gamma(option, stock, model, rate, h_frac, **kw)
        up, mid, dn = self._bumped_prices(option, stock, model, rate,
                                          (1 + h_frac, 1.0, 1 - h_frac), **kw)
        h = stock.spot * h_frac
        return (up - 2 * mid + dn) / (h * h)

    @_engine.serialized
    def rho(self, option, stock, model, rate, h: float = 1e-4, **kw: Any) -> float:
        if not self._scaling_greeks_ok(model):
            return super().rho(option, stock, model, rate, h, **kw)
        up, dn = self._bumped_prices(option, stock, model, rate, (), (h, -h), **kw)
        return (up - dn) / (2 * h)

    @_engine.serialized
    def price_with_greeks(self, option, stock, model, rate, greeks=("delta", "gamma", "vega"),
                          h_frac: float = 1e-3, h_vega: float = 1e-4, h_rho: float = 1e-4,
                          **kw: Any) -> PricingResult:
        """
        Price and greeks in ONE call -> ``PricingResult(price, greeks={...})`` (the ``greeks``
        field the reference's dataclass carries, base/pricing_result.py:10-27, and never fills).
        All numbers are on common random numbers -- the estimators of GreekMixin
        (greek_mixin.py:29-265) -- from as few simulations as the model allows:

        * log-Euler kinds (BSM, Heston, Merton, Bates, Kou): ONE simulation; price, delta, gamma
          and rho are strike rescalings of its terminal values, vega (one-factor kinds) a sigma
          sweep over the (W, J) stored by the same launch;
        * SABR kinds: ONE simulation carrying the spot scenarios (up, base, down) side by side
          for price, delta and gamma; vega is NaN as in the reference (no ``sigma``); rho re-prices;
        * theta: GBM from the Brownian sums the same launch stores (a (drift, sigma) scenario of
          the vega sweep); every other model re-prices (the whole step depends on T);
        * anything else (numpy mode, fp32, single-step models): the mixin's re-pricings.
        The generator advances once, as for one ``price`` call.
        """
        greeks = tuple(greeks)
        unknown = set(greeks) - {"delta", "gamma", "vega", "rho", "theta"}
        if unknown:
            raise ValueError(f"unknown greeks {sorted(unknown)}")
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)
        out: dict[str, float] = {}
        state0 = self.rng.bit_generator.state

        def via_mixin(names):
            for g in names:
                self.rng.bit_generator.state = state0
                fn = getattr(self, g)
                args = {"delta": (h_frac,), "gamma": (h_frac,), "vega": (h_vega,),
                        "rho": (h_rho,), "theta": ()}[g]
                out[g] = float(fn(option, stock, model, rate, *args, **kw))

        if self._scenario_greeks_ok(model):
            self.rng.bit_generator.state = state0
            up, mid, dn = self._scenario_prices(option, stock, model, rate,
                                                (1 + h_frac, 1.0, 1 - h_frac), **kw)
            h = S0 * h_frac
            price = mid
            if "delta" in greeks:
                out["delta"] = (up - dn) / (2 * h)
            if "gamma" in greeks:
                out["gamma"] = (up - 2 * mid + dn) / (h * h)
            if "vega" in greeks and "sigma" not in model.params:
                out["vega"] = float("nan")                    # greek_mixin.py:156-157
            via_mixin([g for g in greeks if g not in out])
        elif self._scaling_greeks_ok(model):
            kind = model_kind(model)
            scen = [(1.0, 0.0), (1 + h_frac, 0.0), (1 - h_frac, 0.0), (1.0, h_rho), (1.0, -h_rho)]
            strikes = [K / (f * math.exp(hh * T)) for f, hh in scen]
            eng = _engine.get_engine()
            num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
            rank, world, all_reduce = self._group()
            _, cnt = shard_draws(num_draws, rank, world)
            want_vega = "vega" in greeks and "sigma" in model.params
            one_sim_vega = want_vega and kind in self._ONE_FACTOR_KINDS
            one_sim_theta = ("theta" in greeks and kind == "bsm" and self.precision == "fp64" and not kw
                             and self._theta_steps_ok(T, 1e-5))
            aux = eng.buffer("aux_wj", 3 * cnt) if (one_sim_vega or one_sim_theta) else None
            self.rng.bit_generator.state = state0
            with crn(self.rng), self._steps_for(T) as steps:
                mom, _, seed = self._sweep(model, S0, r, q, T, strikes, [1 if call else 0] * 5,
                                           _aux=aux, **kw)
                c_mean = control_mean(kind, model.params, S0, r, q, T, steps)
                vals = []
                for (f, hh), m in zip(scen, mom):
                    st = summarize(m, r, T, c_mean)
                    use_cv = self.control_variate and "cv_price" in st and math.isfinite(c_mean)
                    vals.append(f * float(st["cv_price"] if use_cv else st["price"]))
                    if (f, hh) == (1.0, 0.0):
                        st["seed"] = seed
                        self.last_stats = st
                price, up, dn, r_up, r_dn = vals
                h = S0 * h_frac
                if "delta" in greeks:
                    out["delta"] = (up - dn) / (2 * h)
                if "gamma" in greeks:
                    out["gamma"] = (up - 2 * price + dn) / (h * h)
                if "rho" in greeks:
                    out["rho"] = (r_up - r_dn) / (2 * h_rho)
                if one_sim_vega:
                    out["vega"] = self._vega_from_aux(eng, aux, cnt, kind, model, S0, K, T, r, q,
                                                      call, h_vega, steps, all_reduce)
                if one_sim_theta:
                    out["theta"] = self._theta_from_aux(eng, aux, cnt, model, S0, K, T, q, rate, call,
                                                        1e-5, steps, all_reduce)
            if "vega" in greeks and "sigma" not in model.params:
                out["vega"] = float("nan")
            via_mixin([g for g in greeks if g not in out])
        else:
            self.rng.bit_generator.state = state0
            with crn(self.rng):
                price = self.price(option, stock, model, rate, **kw).price
            via_mixin(list(greeks))
        # advance the stream once, as one price() call does
        self.rng.bit_generator.state = state0
        self.rng.integers(0, 2**63 - 1, dtype=np.int64)
        return PricingResult(price=float(price), greeks={g: out[g] for g in greeks})

    def _vega_from_aux(self, eng, aux, cnt, kind, model, S0, K, T, r, q, call, h, steps,
                       all_reduce):
        """Central difference in sigma from the (W, J) one simulation stored (see ``vega``)."""
        p = model.params
        dt = T / steps
        comp = 0.0
        if kind == "merton":
            comp = p["lambda"] * (math.exp(p["mu_j"] + 0.5 * p["sigma_j"] ** 2) - 1)
        elif kind == "kou":
            comp = p["lambda"] * ((p["p_up"] * p["eta1"] / (p["eta1"] - 1))
                                  + ((1 - p["p_up"]) * p["eta2"] / (p["eta2"] + 1)) - 1)
        sigmas = [p["sigma"] + h, p["sigma"] - h]
        a0 = [math.log(S0) + steps * ((r - q - 0.5 * s * s - comp) * dt) for s in sigmas]
        dev = eng.sigma_sweep(aux, cnt, self.antithetic, a0, sigmas, [float(K)] * 2,
                              [1 if call else 0] * 2)
        if all_reduce is not None:
            all_reduce(dev[: 2 * _engine.NMOM])
        mom = eng.read(dev, 2 * _engine.NMOM).reshape(2, _engine.NMOM)
        c_mean = control_mean(kind, p, S0, r, q, T, steps)
        prices = []
        for row in mom:
            st = summarize(row, r, T, c_mean)
            use_cv = self.control_variate and "cv_price" in st and math.isfinite(c_mean)
            prices.append(st["cv_price"] if use_cv else st["price"])
        return float((prices[0] - prices[1]) / (2 * h))

    def _theta_steps_ok(self, T: float, h: float) -> bool:
        """The maturity bumps keep the step count (always, unless ``steps_per_year`` rounds
        T + h or T - h to another grid)."""
        if self.steps_per_year is None:
            return True
        n = [max(1, int(round(self.steps_per_year * t))) for t in (T, T + h, max(T - h, 1e-12))]
        return n[0] == n[1] == n[2]

    def _theta_from_aux(self, eng, aux, cnt, model, S0, K, T, q, rate, call, h, steps, all_reduce):
        """
        GreekMixin.theta (greek_mixin.py:175-218) for the GBM kind from the Brownian sums W of
        ONE simulation: on the same draws a maturity T' has dt' = T'/n, so
        log S_T' = log S_0 + n drift dt' + sigma sqrt(T'/T) W -- a (drift, sigma) scenario of the
        sweep that computes vega.  (Jump kinds: the per-path jump count depends on lambda T', so
        their bumped prices are not functions of the stored (W, J); they re-price.)
        """
        p = model.params
        sig = p["sigma"]
        Ts = [T + h, max(T - h, 1e-12)]
        rs = [rate.get_rate(t) for t in Ts]
        a0 = [math.log(S0) + steps * ((rr - q - 0.5 * sig * sig) * (t / steps)) for t, rr in zip(Ts, rs)]
        dev = eng.sigma_sweep(aux, cnt, self.antithetic, a0, [sig * math.sqrt(t / T) for t in Ts],
                              [float(K)] * 2, [1 if call else 0] * 2)
        if all_reduce is not None:
            all_reduce(dev[: 2 * _engine.NMOM])
        mom = eng.read(dev, 2 * _engine.NMOM).reshape(2, _engine.NMOM)
        prices = []
        for row, t, rr in zip(mom, Ts, rs):
            c_mean = control_mean("bsm", p, S0, rr, q, t, steps)
            st = summarize(row, rr, t, c_mean)
            use_cv = self.control_variate and "cv_price" in st and math.isfinite(c_mean)
            prices.append(st["cv_price"] if use_cv else st["price"])
        return float((prices[1] - prices[0]) / (2 * h))

    @_engine.serialized
    def theta(self, option, stock, model, rate, h: float = 1e-5, **kw: Any) -> float:
        """
        GreekMixin.theta (greek_mixin.py:175-218): central difference in the maturity on common
        random numbers.  GBM on the device RNG: both bumped prices are read off ONE simulation
        that stores the Brownian sums (``_theta_from_aux``); every other model re-prices twice,
        as the reference does.
        """
        kind = model_kind(model) if self._philox_sde(model) else None
        T = option.maturity
        if (kind != "bsm" or self.precision != "fp64" or kw or T <= 0
                or not self._theta_steps_ok(T, h)):
            return super().theta(option, stock, model, rate, h, **kw)
        eng = _engine.get_engine()
        S0, K = stock.spot, option.strike
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        with crn(self.rng):                       # greeks do not advance the stream
            seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        mdl = _engine.make_model(kind, model.params, r, q)
        rank, world, all_reduce = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        with self._steps_for(T) as steps:
            sim = eng.make_sim(float(S0), T, num_draws, steps, self.antithetic, seed,
                               _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
            aux = eng.buffer("aux_wj", 3 * cnt)
            eng.european_async(mdl, sim, [float(K)], [1 if call else 0],
                               eng.buffer("moments", _engine.NMOM), aux=aux)
            return self._theta_from_aux(eng, aux, cnt, model, S0, K, T, q, rate, call, h, steps,
                                        all_reduce)

    # ------------------------------------------- single-step models
    def _device_single_step(self, model) -> bool:
        """VG / NIG / CGMY / Hyperbolic / CEV on the device samplers (csrc/levy.cuh)."""
        return self.rng_mode == "philox" and levy_kind(model) is not None

    def _device_terminal(self, model, S0, r, q, T, hint=None):
        """
        Terminal spots of ``n_paths`` samples of a single-step model drawn on the
        GPU (one thread per sample; samples sharded by rank like the SDE draws).
        With ``hint`` = (strike, is_call) only the payoff moments come back
        (DeviceTerminal); without it the host array of terminal spots.
        """
        kind = levy_kind(model)
        eng = _engine.get_engine()
        levy = _engine.make_levy(kind, model.params, r, T, S0)
        if kind == "cev":
            drift = 0.0                                   # cev.py:49-96 samples S_T directly (q unused)
        else:
            compensator = float(np.log(np.real(model.raw_cf(t=T)(-1j))))     # monte_carlo.py:276-278
            drift = (r - q) * T - compensator
        seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        rank, world, all_reduce = self._group()
        off, cnt = shard_draws(self.n_paths, rank, world)
        dev = eng.buffer("moments", _engine.NMOM)
        if hint is None:
            term = eng.buffer("terminal", cnt)
            eng.levy_terminal_async(levy, S0, drift, seed, off, cnt, [float(S0)], [1], dev, term)
            return term[:cnt].cpu().numpy()
        K, call = hint
        eng.levy_terminal_async(levy, S0, drift, seed, off, cnt, [K], [1 if call else 0], dev)
        if all_reduce is not None:
            all_reduce(dev[: _engine.NMOM])
        return DeviceTerminal(eng.read(dev, _engine.NMOM), kind, seed)

    _ONE_FACTOR_KINDS = ("bsm", "merton", "kou")

    @_engine.serialized
    def vega(self, option, stock, model, rate, h: float = 1e-4, **kw: Any) -> float:
        """
        GreekMixin.vega (greek_mixin.py:125-175): central difference in ``sigma`` on common
        random numbers.  For the one-factor log-Euler models log S_T = log S_0 + n drift(sigma)
        + sigma W + J with W, J independent of sigma, so both bumped prices are read off ONE
        simulation that stores (W, J) per path; other models re-price twice as the reference.
        """
        if "sigma" not in model.params:
            return np.nan
        kind = model_kind(model) if self._philox_sde(model) else None
        if kind not in self._ONE_FACTOR_KINDS or self.precision != "fp64":
            return super().vega(option, stock, model, rate, h, **kw)
        eng = _engine.get_engine()
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        with crn(self.rng):                       # greeks do not advance the stream
            seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        mdl = _engine.make_model(kind, model.params, r, q)
        rank, world, all_reduce = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed,
                           _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
        aux = eng.buffer("aux_wj", 3 * cnt)
        eng.european_async(mdl, sim, [float(K)], [1 if call else 0],
                           eng.buffer("moments", _engine.NMOM), aux=aux)
        p = model.params
        dt = T / self.n_steps
        comp = 0.0
        if kind == "merton":
            comp = p["lambda"] * (math.exp(p["mu_j"] + 0.5 * p["sigma_j"] ** 2) - 1)
        elif kind == "kou":
            comp = p["lambda"] * ((p["p_up"] * p["eta1"] / (p["eta1"] - 1))
                                  + ((1 - p["p_up"]) * p["eta2"] / (p["eta2"] + 1)) - 1)
        sigmas = [p["sigma"] + h, p["sigma"] - h]
        a0 = [math.log(S0) + self.n_steps * ((r - q - 0.5 * s * s - comp) * dt) for s in sigmas]
        dev = eng.sigma_sweep(aux, cnt, self.antithetic, a0, sigmas, [float(K)] * 2,
                              [1 if call else 0] * 2)
        if all_reduce is not None:
            all_reduce(dev[: 2 * _engine.NMOM])
        mom = eng.read(dev, 2 * _engine.NMOM).reshape(2, _engine.NMOM)
        c_mean = control_mean(kind, p, S0, r, q, T, self.n_steps)
        prices = []
        for row in mom:
            st = summarize(row, r, T, c_mean)
            prices.append(st["cv_price"] if self.control_variate and "cv_price" in st else st["price"])
        return (prices[0] - prices[1]) / (2 * h)

    @_engine.serialized
    def implied_volatility(self, option, stock, model, rate, target_price: float,
                           low: float = 1e-6, high: float = 5.0, tol: float = 1e-6,
                           **kwargs: Any) -> float:
        """
        IVMixin.implied_volatility (iv_mixin.py:28-116): the Black-Scholes volatility at
        which THIS technique reproduces ``target_price``.  The reference re-simulates at
        every trial volatility with fresh draws, so its objective is noisy; here ONE
        simulation stores the Brownian sums W and every trial sigma is a sweep over them
        (log S_T = log S_0 + n drift(sigma) + sigma W): a smooth, monotone objective on
        common random numbers, same brentq bracket and tolerance.  Falls back to the
        mixin (re-simulation) outside the fp64 device-RNG mode.
        """
        if self.rng_mode != "philox" or self.precision != "fp64":
            return super().implied_volatility(option, stock, model, rate, target_price, low, high,
                                              tol, **kwargs)
        from scipy.optimize import brentq

        eng = _engine.get_engine()
        S0, K, T = stock.spot, option.strike, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        call = _is_call(option)
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))     # one draw, like one price()
        mdl = _engine.make_model("bsm", {"sigma": 0.3}, r, q)
        rank, world, all_reduce = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed,
                           _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
        aux = eng.buffer("aux_wj", 3 * cnt)
        eng.european_async(mdl, sim, [float(K)], [1 if call else 0],
                           eng.buffer("moments", _engine.NMOM), aux=aux)
        dt, disc = T / self.n_steps, math.exp(-r * T)

        def gap(vol: float) -> float:
            a0 = math.log(S0) + self.n_steps * ((r - q - 0.5 * vol * vol) * dt)
            dev = eng.sigma_sweep(aux, cnt, self.antithetic, [a0], [vol], [float(K)],
                                  [1 if call else 0])
            if all_reduce is not None:
                all_reduce(dev[: _engine.NMOM])
            m = eng.read(dev, _engine.NMOM)
            p = disc * m[1] / m[0]
            return p - target_price if np.isfinite(p) else 1e6

        try:
            return brentq(gap, low, high, xtol=tol, disp=False)
        except (ValueError, RuntimeError):
            pass
        try:
            return self._secant_iv(gap, 0.2, tol, 100)
        except (ValueError, RuntimeError):
            return np.nan

    def _simulate_levy_terminal(self, model, S0: float, r: float, q: float, T: float):
        """
        Pure-Levy terminal sampling (monte_carlo.py:259-280).  The four laws the
        reference ships (VG, NIG, CGMY with Y=1, Hyperbolic) are drawn on the GPU;
        a model that brings its own ``sample_terminal_log_return`` (the reference's
        plug-in contract, and ``rng_mode="numpy"``) has that sampler called exactly
        as the reference does, with the martingale correction applied here.
        """
        if self._device_single_step(model):
            if not hasattr(model, "raw_cf"):
                raise NotImplementedError(f"Levy model '{model.name}' is missing required methods.")
            return self._device_terminal(model, S0, r, q, T, hint=self._hint)
        if self.rng_mode == "numpy" and not hasattr(model, "sample_terminal_log_return") \
                and levy_kind(model) is not None:
            raise NotImplementedError(
                f"rng_mode='numpy' prices '{model.name}' through the model's own host sampler "
                "(sample_terminal_log_return, as the reference's model classes have); this "
                "engine's model classes sample on the device only -- use rng_mode='philox' or "
                "pass the reference's model object")
        if not hasattr(model, "sample_terminal_log_return") or not hasattr(model, "raw_cf"):
            raise NotImplementedError(f"Levy model '{model.name}' is missing required methods.")
        x = model.sample_terminal_log_return(T, self.n_paths, self.rng)
        compensator = np.log(np.real(model.raw_cf(t=T)(-1j)))
        return S0 * np.exp((r - q) * T - compensator + x)
