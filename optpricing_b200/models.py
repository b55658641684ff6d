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
