"""
AmericanMonteCarloTechnique on the B200 engine: a drop-in for
/root/reference/src/optpricing/techniques/american_monte_carlo.py.

Same constructor, ``price`` signature, error message and private-method surface
(american_monte_carlo.py:35-207).  In the default ``rng_mode="philox"`` the two
stages of the reference -- simulate the full path matrix, then run
Longstaff-Schwartz over it -- still exist, but what passes between them is a
handle to a step-major path store in HBM instead of an (n_paths, n_steps+1)
host matrix.  ``rng_mode="numpy"`` replays the reference's host draws through the
GPU fed-draw path kernels and the same GPU induction (prices match the
reference to ~1e-12 for the same seed).

As in the reference, the two Brownian drivers are drawn independently here
(american_monte_carlo.py:181-185): there is no European-style pre-correlation.
"""
from __future__ import annotations

import math
from typing import Any

import numpy as np

from optpricing_b200 import _engine
from optpricing_b200.atoms import is_call as _is_call
from optpricing_b200.models import model_kind

from .base import BaseTechnique, GreekMixin, IVMixin, PricingResult
from .kernels import path_kernels
from .kernels.american_mc_kernels import DevicePathStore, longstaff_schwartz_pricer
from .monte_carlo import default_rng_mode, shard_draws

_PATH_KERNELS = {
    "bsm": path_kernels.bsm_path_kernel, "heston": path_kernels.heston_path_kernel,
    "merton": path_kernels.merton_path_kernel, "bates": path_kernels.bates_path_kernel,
    "sabr": path_kernels.sabr_path_kernel, "sabr_jump": path_kernels.sabr_jump_path_kernel,
    "kou": path_kernels.kou_path_kernel,
}


class AmericanMonteCarloTechnique(BaseTechnique, GreekMixin, IVMixin):
    def __init__(self, *, n_paths: int = 20_000, n_steps: int = 100, antithetic: bool = True,
                 seed: int | None = None, lsm_degree: int = 2, rng_mode: str | None = None,
                 distributed: bool | None = None, exchange: str = "nvlink"):
        if antithetic and n_paths % 2 != 0:
            n_paths += 1                       # american_monte_carlo.py:44-45
        rng_mode = rng_mode or default_rng_mode()
        if rng_mode not in ("philox", "numpy"):
            raise ValueError("rng_mode must be 'philox' or 'numpy'")
        if exchange not in ("nvlink", "nccl"):
            raise ValueError("exchange must be 'nvlink' or 'nccl'")
        self.n_paths = n_paths
        self.n_steps = n_steps
        self.antithetic = antithetic
        self.rng = np.random.default_rng(seed)
        self.lsm_degree = lsm_degree
        self.exchange = exchange
        self.rng_mode = rng_mode
        self.distributed = distributed

    @_engine.serialized
    def price(self, option, stock, model, rate, **kwargs: Any) -> PricingResult:
        if not model.supports_sde:
            raise TypeError("American MC pricing requires an SDE model.")
        spot_paths = self._simulate_full_sde_paths(option, stock, model, rate, _handle=True,
                                                   **kwargs)
        _, world, all_reduce = self._group()
        sharded = world > 1 and isinstance(spot_paths, DevicePathStore)
        # Several ranks: the per-date Gram exchange runs inside the persistent kernel
        # over NVLink peer memory (default), or as one NCCL all-reduce per date.
        peers = sharded and self.exchange == "nvlink" and self._nvlink_ready(spot_paths.engine)
        args = dict(stock_paths=spot_paths, K=option.strike, dt=option.maturity / self.n_steps,
                    r=rate.get_rate(option.maturity), is_call=_is_call(option),
                    degree=self.lsm_degree)
        if not sharded:
            return PricingResult(price=longstaff_schwartz_pricer(**args))
        # A joint launch waits, on the device, for the kernels of the other PROCESSES: every
        # rank must actually enter it.  Shard sizes are a function of (n_paths, world) alone, so
        # every rank checks ALL shards locally and all raise together, before anything is
        # launched, when one of them would be empty.  Should a rank still fail to arrive (host
        # error, lost process), the others give up after a wall-clock time-out and poison what
        # they publish (csrc/lsmc.cuh), so every rank that did launch raises EngineError.
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        if min(shard_draws(num_draws, rk, world)[1] for rk in range(world)) < 1:
            raise ValueError(f"n_paths = {self.n_paths} leaves a rank without paths when sharded "
                             f"over {world} ranks; use distributed=False for a job this small")
        value = longstaff_schwartz_pricer(**args, peers=peers,
                                          all_reduce=None if peers else all_reduce)
        return PricingResult(price=value)

    @staticmethod
    def _nvlink_ready(eng) -> bool:
        """Attach the ranks' exchange buffers (once per group); NCCL groups only."""
        import torch.distributed as dist

        if dist.get_backend() != "nccl":
            return False
        try:
            eng.lsmc_peer_attach()
        except _engine.EngineError as exc:      # agreed on by all ranks (see lsmc_peer_attach)
            import warnings
            warnings.warn(f"NVLink exchange unavailable ({exc}); using one NCCL all-reduce per "
                          "exercise date", RuntimeWarning, stacklevel=3)
            return False
        return True

    def _group(self):
        if self.distributed is False:
            return 0, 1, None
        try:
            import torch.distributed as dist
        except ImportError:  # pragma: no cover
            return 0, 1, None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0, 1, None
        return dist.get_rank(), dist.get_world_size(), dist.all_reduce   # per-date sums: NCCL path

    def _get_path_kernel_and_params(self, model, r: float, q: float, dt: float, **kwargs: Any):
        """(path kernel, keyword arguments) as american_monte_carlo.py:75-150 pairs them."""
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
        else:
            raise TypeError("American MC pricing requires an SDE model.")
        if kind in ("sabr_jump", "bates", "merton"):
            kp.update(lambda_=p["lambda"], mu_j=p["mu_j"], sigma_j=p["sigma_j"])
        return _PATH_KERNELS[kind], kp

    def _simulate_full_sde_paths(self, option, stock, model, rate, _handle: bool = False,
                                 **kwargs: Any):
        """
        The (n_paths, n_steps + 1) spot matrix of american_monte_carlo.py:152-207.
        ``price`` asks for a DevicePathStore handle instead (Philox mode), so the
        paths stay in HBM between simulation and induction.
        """
        S0, T = stock.spot, option.maturity
        r, q = rate.get_rate(T), stock.dividend
        if self.rng_mode == "numpy":
            return self._simulate_full_sde_paths_numpy(model, S0, r, q, T, **kwargs)
        kind = model_kind(model)
        if kind not in _PATH_KERNELS:                 # american_monte_carlo.py:149-150
            raise TypeError("American MC pricing requires an SDE model.")
        eng = _engine.get_engine()
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        seed = int(self.rng.integers(0, 2**63 - 1, dtype=np.int64))
        mdl = _engine.make_model(kind, model.params, r, q, v0=kwargs.get("v0"))
        rank, world, _ = self._group()
        off, cnt = shard_draws(num_draws, rank, world)
        n_local_paths = cnt * (2 if self.antithetic else 1)
        sim = eng.make_sim(float(S0), T, num_draws, self.n_steps, self.antithetic, seed,
                           _engine.CORR_INDEPENDENT_DRAWS, off, cnt)
        store, ld = eng.lsmc_store(n_local_paths, self.n_steps)
        eng.lsmc_fill_paths(mdl, sim, store, ld)
        handle = DevicePathStore(eng, store, ld, n_local_paths, self.n_steps, float(S0))
        return handle if _handle else np.asarray(handle)

    def _simulate_full_sde_paths_numpy(self, model, S0, r, q, T, **kwargs):
        dt = T / self.n_steps
        num_draws = self.n_paths // 2 if self.antithetic else self.n_paths
        kernel, kp = self._get_path_kernel_and_params(model, r, q, dt, **kwargs)
        sabr = getattr(model, "is_sabr", False)
        args = {"n_paths": num_draws, "n_steps": self.n_steps, **kp}
        args["s0" if sabr else "log_s0"] = S0 if sabr else math.log(S0)
        shape, sd = (num_draws, self.n_steps), math.sqrt(dt)
        if model.has_variance_process:
            z1 = self.rng.standard_normal(size=shape)
            z2 = self.rng.standard_normal(size=shape)
            args["dw1"], args["dw2"] = z1 * sd, z2 * sd
        else:
            args["dw"] = self.rng.standard_normal(size=shape) * sd
        if model.has_jumps:
            args["jump_counts"] = self.rng.poisson(lam=model.params["lambda"] * dt, size=shape)
        paths = kernel(**args)
        if self.antithetic:
            for key in ("dw", "dw1", "dw2"):
                if key in args:
                    args[key] = -args[key]
            paths = np.concatenate([paths, kernel(**args)])
        return paths
