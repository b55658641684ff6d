"""B200-backed Monte Carlo techniques with the reference's API (techniques/__init__.py:8-15)."""
from .american_monte_carlo import AmericanMonteCarloTechnique
from .base import BaseTechnique, GreekMixin, IVMixin, PricingResult, crn
from .monte_carlo import MonteCarloTechnique

__all__ = ["BaseTechnique", "PricingResult", "GreekMixin", "IVMixin", "crn",
           "MonteCarloTechnique", "AmericanMonteCarloTechnique"]
