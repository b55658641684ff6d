"""
optpricing_b200 -- a B200-native (sm_100a) Monte Carlo engine behind the API of
Diljit22/optpricing's MonteCarloTechnique and AmericanMonteCarloTechnique.

    from optpricing_b200 import (MonteCarloTechnique, AmericanMonteCarloTechnique,
                                 Option, OptionType, Stock, Rate, HestonModel)
    MonteCarloTechnique(n_paths=1 << 24, n_steps=252, seed=42).price(option, stock, model, rate)

``optpricing_b200.overlay.install()`` makes the reference package itself resolve
its Monte Carlo modules to this engine.

Importing the package needs neither a GPU nor the built library; pricing does,
and fails loudly without them (there is no CPU fallback).
"""
from .atoms import ExerciseStyle, Option, OptionType, Rate, Stock
from .models import (BatesModel, BSMModel, CEVModel, CGMYModel, DupireLocalVolModel, HestonModel,
                     HyperbolicModel,
                     KouModel, MertonJumpModel, NIGModel, SABRJumpModel, SABRModel,
                     VarianceGammaModel, levy_kind, model_kind)
from .techniques import (AmericanMonteCarloTechnique, BaseTechnique, GreekMixin, IVMixin,
                         MonteCarloTechnique, PricingResult, crn)

__version__ = "0.1.0"

__all__ = [
    "Option", "OptionType", "ExerciseStyle", "Stock", "Rate",
    "BSMModel", "HestonModel", "MertonJumpModel", "BatesModel", "KouModel", "SABRModel",
    "SABRJumpModel", "DupireLocalVolModel", "VarianceGammaModel", "NIGModel", "CGMYModel", "HyperbolicModel", "CEVModel",
    "model_kind", "levy_kind",
    "MonteCarloTechnique", "AmericanMonteCarloTechnique", "BaseTechnique", "GreekMixin",
    "IVMixin", "PricingResult", "crn",
]
