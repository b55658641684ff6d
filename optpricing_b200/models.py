"""
Parameter containers for the seven SDE models on the Monte Carlo path.

Mirrors only what the techniques read from the reference's model classes
(/root/reference/src/optpricing/models/{bsm,heston,merton_jump,bates,kou,sabr,
sabr_jump}.py): ``name``, ``params``, ``default_params``, the capability flags
(models/base/base_model.py:47-55) and ``with_params``.  Characteristic
functions, closed forms and PDE coefficients are out of scope.
The reference's own model objects are accepted wherever these are.
"""
from __future__ import annotations


class SDEModel:
    name = "SDEModel"
    supports_sde = True
    supports_cf = False
    has_closed_form = False
    has_variance_process = False
    has_jumps = False
    is_pure_levy = False
    is_sabr = False
    default_params: dict = {}
    required: tuple = ()
    positive: tuple = ()
    bounded: tuple = ()   # (key, low, high)

    def __init__(self, params: dict | None = None):
        self.params = dict(params) if params else dict(self.default_params)
        missing = [k for k in self.required if k not in self.params]
        if missing:
            raise ValueError(f"{self.name}: missing required parameters {missing}")
        for k in self.positive:
            if not self.params[k] > 0:
                raise ValueError(f"{self.name}: parameter '{k}' must be positive")
        for k, lo, hi in self.bounded:
            if not lo <= self.params[k] <= hi:
                raise ValueError(f"{self.name}: parameter '{k}' must be in [{lo}, {hi}]")

    def with_params(self, **updated):
        return type(self)(params={**self.params, **updated})

    def __repr__(self):
        return f"{type(self).__name__}({self.params})"

    def __eq__(self, other):
        return type(other) is type(self) and other.params == self.params

    def __hash__(self):
        return hash((type(self), tuple(sorted(self.params.items()))))


class BSMModel(SDEModel):
    name = "Black-Scholes-Merton"
    has_closed_form = True
    default_params = {"sigma": 0.2}
    required = positive = ("sigma",)


class HestonModel(SDEModel):
    name = "Heston"
    has_variance_process = True
    default_params = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5}
    required = ("kappa", "theta", "rho", "vol_of_vol")
    positive = ("kappa", "theta", "vol_of_vol")
    bounded = (("rho", -1.0, 1.0),)


class MertonJumpModel(SDEModel):
    name = "Merton Jump-Diffusion"
    has_jumps = True
    default_params = {"sigma": 0.2, "lambda": 0.5, "mu_j": -0.1, "sigma_j": 0.15,
                      "max_sum_terms": 100}
    required = ("sigma", "lambda", "mu_j", "sigma_j")
    positive = ("sigma", "sigma_j")


class BatesModel(SDEModel):
    name = "Bates"
    has_variance_process = True
    has_jumps = True
    default_params = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5,
                      "lambda": 0.5, "mu_j": -0.1, "sigma_j": 0.15}
    required = ("kappa", "theta", "rho", "vol_of_vol", "lambda", "mu_j", "sigma_j")
    positive = ("kappa", "theta", "vol_of_vol", "sigma_j")
    bounded = (("rho", -1.0, 1.0),)


class KouModel(SDEModel):
    name = "Kou Double-Exponential Jump"
    has_jumps = True
    default_params = {"sigma": 0.15, "lambda": 1.0, "p_up": 0.6, "eta1": 10.0, "eta2": 5.0}
    required = ("sigma", "lambda", "p_up", "eta1", "eta2")
    positive = ("sigma", "lambda", "eta1", "eta2")
    bounded = (("p_up", 0.0, 1.0),)


class SABRModel(SDEModel):
    name = "SABR"
    has_variance_process = True
    is_sabr = True
    default_params = {"alpha": 0.5, "beta": 0.8, "rho": -0.6}
    required = ("alpha", "beta", "rho")
    positive = ("alpha",)
    bounded = (("beta", 0.0, 1.0), ("rho", -1.0, 1.0))


class SABRJumpModel(SDEModel):
    name = "SABR with Jumps"
    has_variance_process = True
    has_jumps = True
    is_sabr = True
    default_params = {"alpha": 0.5, "beta": 0.8, "rho": -0.6, "lambda": 0.4, "mu_j": -0.1,
                      "sigma_j": 0.15}
    required = ("alpha", "beta", "rho", "lambda", "mu_j", "sigma_j")


# ---------------------------------------------------------------------------
# Single-step models: the pure-Levy and exact-sampler branches of
# MonteCarloTechnique.price (monte_carlo.py:103-106).  Only what the technique
# reads: the flags, ``params`` and ``raw_cf`` (the martingale correction of
# monte_carlo.py:276-278 is log Re phi_raw(-i), a host scalar).  The samplers
# themselves (sample_terminal_log_return / sample_terminal_spot in the
# reference) run on the device: csrc/levy.cuh.
# ---------------------------------------------------------------------------
class _LevyModel(SDEModel):
    supports_sde = False
    supports_cf = True
    is_pure_levy = True


class VarianceGammaModel(_LevyModel):
    name = "Variance Gamma"
    default_params = {"sigma": 0.2, "nu": 0.1, "theta": -0.14}
    required = ("sigma", "nu", "theta")
    positive = ("sigma", "nu")

    def raw_cf(self, *, t: float):                      # models/vg.py:111-134
        s, nu, th = self.params["sigma"], self.params["nu"], self.params["theta"]
        return lambda u: (1 - 1j * u * th * nu + 0.5 * u ** 2 * s ** 2 * nu) ** (-t / nu)


class NIGModel(_LevyModel):
    name = "Normal Inverse Gaussian"
    default_params = {"alpha": 15.0, "beta": -5.0, "delta": 0.5}
    required = ("alpha", "beta", "delta")
    positive = ("alpha", "delta")

    def __init__(self, params: dict | None = None):
        super().__init__(params)
        if not abs(self.params["beta"]) < self.params["alpha"]:
            raise ValueError("NIG params must satisfy |beta| < alpha.")

    def raw_cf(self, *, t: float):                      # models/nig.py:103-134
        import numpy as np
        a, b, d = self.params["alpha"], self.params["beta"], self.params["delta"]
        return lambda u: np.exp(d * t * (np.sqrt(a ** 2 - b ** 2)
                                         - np.sqrt(a ** 2 - (b + 1j * u) ** 2)))


class CGMYModel(_LevyModel):
    name = "CGMY"
    default_params = {"C": 0.02, "G": 5.0, "M": 5.0, "Y": 1.2}
    required = ("C", "G", "M", "Y")
    positive = ("C", "G", "M")

    def __init__(self, params: dict | None = None):
        super().__init__(params)
        if not self.params["Y"] < 2:
            raise ValueError("CGMY: parameter 'Y' must be less than 2")

    def raw_cf(self, *, t: float):                      # models/cgmy.py:103-132
        import numpy as np
        from scipy.special import gamma as gamma_func     # the reference's: inf at Y = 1 -> nan
        C, G, M, Y = (self.params[k] for k in ("C", "G", "M", "Y"))
        return lambda u: np.exp(C * t * gamma_func(-Y) * ((M - 1j * u) ** Y - M ** Y
                                                          + (G + 1j * u) ** Y - G ** Y))


class HyperbolicModel(_LevyModel):
    name = "Hyperbolic"
    default_params = {"alpha": 15.0, "beta": -5.0, "delta": 0.5, "mu": 0.0}
    required = ("alpha", "beta", "delta", "mu")
    positive = ("alpha", "delta")

    def __init__(self, params: dict | None = None):
        super().__init__(params)
        if not abs(self.params["beta"]) < self.params["alpha"]:
            raise ValueError("Hyperbolic params must satisfy |beta| < alpha.")

    def raw_cf(self, *, t: float):                      # models/hyperbolic.py:92-121
        import numpy as np
        from scipy.special import kv
        a, b, d, mu = (self.params[k] for k in ("alpha", "beta", "delta", "mu"))
        g0 = np.sqrt(a ** 2 - b ** 2)

        def phi_raw(u):
            gu = np.sqrt(a ** 2 - (b + 1j * u) ** 2)
            return np.exp(1j * u * mu * t) * (g0 / gu) ** t * (kv(1, d * t * gu) / kv(1, d * t * g0))
        return phi_raw


def device_sample_terminal_spot(self, S0: float, r: float, T: float, size: int):
    """
    CEVModel.sample_terminal_spot (models/cev.py:49-96) on the GPU: ``size``
    terminal spots as a host array.  Like the reference it takes no Generator;
    the Philox key is one draw from numpy's global state, so ``np.random.seed``
    makes it reproducible.  ``MonteCarloTechnique.price`` recognises this
    function and keeps the samples on the device (only the payoff moments
    come back); a patched or user-supplied sampler is called as is.
    """
    import numpy as np

    from . import _engine
    eng = _engine.get_engine()
    levy = _engine.make_levy("cev", self.params, r, T, S0)
    seed = int(np.random.randint(0, 2**31 - 1)) | (int(np.random.randint(0, 2**31 - 1)) << 31)  # noqa: NPY002
    term = eng.buffer("terminal", int(size))
    mom = eng.buffer("moments", _engine.NMOM)
    eng.levy_terminal_async(levy, S0, 0.0, seed, 0, int(size), [float(S0)], [1], mom, term)
    return term[: int(size)].cpu().numpy()


device_sample_terminal_spot.__optmc_device__ = True


class CEVModel(SDEModel):
    name = "CEV"
    supports_sde = True
    has_exact_sampler = True
    default_params = {"sigma": 0.8, "gamma": 0.7}
    required = ("sigma", "gamma")
    positive = ("sigma",)
    sample_terminal_spot = device_sample_terminal_spot


_LEVY_KINDS = {"Variance Gamma": "vg", "Normal Inverse Gaussian": "nig", "CGMY": "cgmy",
               "Hyperbolic": "hyperbolic", "CEV": "cev"}


def levy_kind(model) -> str | None:
    """Device sampler name of a single-step model (ours or the reference's), else None."""
    return _LEVY_KINDS.get(getattr(model, "name", None))


class DupireLocalVolModel(SDEModel):
    """models/dupire_local.py:13-43: params = {"vol_surface": callable(T, K) -> local vol}."""
    name = "Dupire Local Volatility"
    is_local_vol = True
    default_params = {"vol_surface": lambda T, K: 0.2}
    required = ("vol_surface",)

    def __init__(self, params: dict | None = None):
        super().__init__(params)
        if not callable(self.params["vol_surface"]):
            raise ValueError("Dupire model requires a callable 'vol_surface' in its parameters.")

    def __repr__(self):
        f = self.params["vol_surface"]
        return f"{type(self).__name__}(vol_surface={getattr(f, '__name__', str(type(f)))})"


def model_kind(model) -> str:
    """
    The reference's dispatch chain (monte_carlo.py:130-197) as a kernel name.
    Duck-typed on the capability flags, so reference model objects work too.
    """
    if getattr(model, "is_sabr", False):
        return "sabr_jump" if model.has_jumps else "sabr"
    if getattr(model, "is_local_vol", False):
        return "dupire"
    if model.has_variance_process and model.has_jumps:
        return "bates"
    if model.has_variance_process:
        return "heston"
    if model.name == "Kou Double-Exponential Jump":
        return "kou"
    if model.has_jumps:
        return "merton"
    return "bsm"
