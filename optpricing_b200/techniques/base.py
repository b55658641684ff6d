"""
Host-side mirror of the reference's technique base layer
(/root/reference/src/optpricing/techniques/base/): the ``price`` ABC, the
``PricingResult`` container, the common-random-number context manager and the
bump-and-reprice greeks / implied-vol mixins.  Signatures, defaults, bump sizes
and return types follow the reference so that callers can switch packages:

  base_technique.py:16-54   BaseTechnique.price(option, stock, model, rate, **kw)
  pricing_result.py:10-27   PricingResult(price, greeks={}, implied_vol=None)
  random_utils.py:12-31     crn(rng)
  greek_mixin.py:29-265     delta/gamma (h = spot*1e-3), vega (h=1e-4, NaN without
                            'sigma'), theta (h=1e-5), rho (h=1e-4)
  iv_mixin.py:28-116        brentq on a BSM re-pricing, secant fallback

The greeks are expressed through one helper that evaluates ``self.price`` on a
list of bumped inputs, each inside ``crn`` when the technique owns a numpy
Generator -- for the Monte Carlo techniques that replays the same Philox key, so
the bumped prices share every draw.
"""
from __future__ import annotations

import abc
import contextlib
import dataclasses
from typing import Any

import numpy as np


@dataclasses.dataclass(frozen=True)
class PricingResult:
    price: float
    greeks: dict = dataclasses.field(default_factory=dict)
    implied_vol: float | None = None


class BaseTechnique(abc.ABC):
    @abc.abstractmethod
    def price(self, option, stock, model, rate, **kwargs: Any) -> PricingResult:
        raise NotImplementedError


@contextlib.contextmanager
def crn(rng: np.random.Generator):
    """Replay the generator: whatever is drawn inside is drawn again next time."""
    saved = rng.bit_generator.state
    try:
        yield
    finally:
        rng.bit_generator.state = saved


class GreekMixin:
    def _reprice(self, scenarios, **kw):
        """Prices for [(option, stock, model, rate), ...] on common random numbers."""
        rng = getattr(self, "rng", None)
        out = []
        for option, stock, model, rate in scenarios:
            guard = crn(rng) if isinstance(rng, np.random.Generator) else contextlib.nullcontext()
            with guard:
                out.append(self.price(option, stock, model, rate, **kw).price)
        return out

    def delta(self, option, stock, model, rate, h_frac: float = 1e-3, **kwargs: Any) -> float:
        h = stock.spot * h_frac
        up, dn = self._reprice(
            [(option, dataclasses.replace(stock, spot=stock.spot + h), model, rate),
             (option, dataclasses.replace(stock, spot=stock.spot - h), model, rate)], **kwargs)
        return (up - dn) / (2 * h)

    def gamma(self, option, stock, model, rate, h_frac: float = 1e-3, **kw: Any) -> float:
        h = stock.spot * h_frac
        up, mid, dn = self._reprice(
            [(option, dataclasses.replace(stock, spot=stock.spot + h), model, rate),
             (option, stock, model, rate),
             (option, dataclasses.replace(stock, spot=stock.spot - h), model, rate)], **kw)
        return (up - 2 * mid + dn) / (h * h)

    def vega(self, option, stock, model, rate, h: float = 1e-4, **kw: Any) -> float:
        if "sigma" not in model.params:
            return np.nan
        sigma = model.params["sigma"]
        up, dn = self._reprice(
            [(option, stock, model.with_params(sigma=sigma + h), rate),
             (option, stock, model.with_params(sigma=sigma - h), rate)], **kw)
        return (up - dn) / (2 * h)

    def theta(self, option, stock, model, rate, h: float = 1e-5, **kw: Any) -> float:
        T0 = option.maturity
        later, sooner = self._reprice(
            [(dataclasses.replace(option, maturity=T0 + h), stock, model, rate),
             (dataclasses.replace(option, maturity=max(T0 - h, 1e-12)), stock, model, rate)],
            **kw)
        return (sooner - later) / (2 * h)

    def rho(self, option, stock, model, rate, h: float = 1e-4, **kw: Any) -> float:
        r0 = rate.get_rate(option.maturity)
        up, dn = self._reprice(
            [(option, stock, model, dataclasses.replace(rate, rate=r0 + h)),
             (option, stock, model, dataclasses.replace(rate, rate=r0 - h))], **kw)
        return (up - dn) / (2 * h)


class IVMixin:
    def implied_volatility(self, option, stock, model, rate, target_price: float,
                           low: float = 1e-6, high: float = 5.0, tol: float = 1e-6,
                           **kwargs: Any) -> float:
        from scipy.optimize import brentq

        from optpricing_b200.models import BSMModel

        probe = BSMModel(params={"sigma": 0.3})

        def gap(vol: float) -> float:
            try:
                with np.errstate(all="ignore"):
                    p = self.price(option, stock, probe.with_params(sigma=vol), rate).price
            except (ZeroDivisionError, OverflowError):
                return 1e6
            return p - target_price if np.isfinite(p) else 1e6

        try:
            return brentq(gap, low, high, xtol=tol, disp=False)
        except (ValueError, RuntimeError):
            pass
        try:
            return self._secant_iv(gap, 0.2, tol, 100)
        except (ValueError, RuntimeError):
            return np.nan

    @staticmethod
    def _secant_iv(fn, x0: float, tol: float, max_iter: int) -> float:
        xa, xb = x0, x0 * 1.1
        fa = fn(xa)
        for _ in range(max_iter):
            fb = fn(xb)
            if abs(fb) < tol:
                return xb
            slope = fb - fa
            if abs(slope) < 1e-14:
                break
            xa, xb, fa = xb, xb - fb * (xb - xa) / slope, fb
        if abs(fn(xb)) < tol * 10:
            return xb
        raise RuntimeError("Secant method failed to converge.")
