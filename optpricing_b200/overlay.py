"""
Drop-in installation: make the reference package resolve its Monte Carlo
modules to this engine, leaving everything else (models, closed forms, FFT,
lattices, calibration, CLI, dashboard) untouched.

    import optpricing_b200.overlay as overlay
    overlay.install()                      # before or after `import optpricing`
    from optpricing.techniques import MonteCarloTechnique      # -> B200 engine

What is replaced:
    optpricing.techniques.monte_carlo                 -> optpricing_b200.techniques.monte_carlo
    optpricing.techniques.american_monte_carlo        -> ....american_monte_carlo
    optpricing.techniques.kernels.mc_kernels          -> ....kernels.mc_kernels
    optpricing.techniques.kernels.path_kernels        -> ....kernels.path_kernels
    optpricing.techniques.kernels.american_mc_kernels -> ....kernels.american_mc_kernels
    optpricing.models.cev.CEVModel.sample_terminal_spot -> device sampler (csrc/levy.cuh)
The aliases are entries in ``sys.modules``, so ``from .monte_carlo import
MonteCarloTechnique`` inside the reference's own ``techniques/__init__.py`` (:15),
``unittest.mock.patch("optpricing.techniques.monte_carlo....")`` and the
``kernel_func is mc_kernels.bsm_kernel`` identity checks of the reference's tests
all land on this engine's objects.

If ``optpricing`` is not importable as a distribution (a source checkout whose
top-level ``__init__`` needs package metadata and yfinance), pass
``reference_src=".../src"`` (or set OPTPRICING_REFERENCE_SRC) and a namespace
shim is created so that sub-packages import without the top-level ``__init__``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_ALIASES = {
    "optpricing.techniques.monte_carlo": "optpricing_b200.techniques.monte_carlo",
    "optpricing.techniques.american_monte_carlo": "optpricing_b200.techniques.american_monte_carlo",
    "optpricing.techniques.kernels.mc_kernels": "optpricing_b200.techniques.kernels.mc_kernels",
    "optpricing.techniques.kernels.path_kernels": "optpricing_b200.techniques.kernels.path_kernels",
    "optpricing.techniques.kernels.american_mc_kernels":
        "optpricing_b200.techniques.kernels.american_mc_kernels",
}
_CLASSES = ("MonteCarloTechnique", "AmericanMonteCarloTechnique")


def install(reference_src: str | None = None, rng_mode: str | None = None) -> None:
    """``rng_mode`` ("philox" | "numpy" | None): the draw source of techniques constructed
    without an explicit ``rng_mode=`` from now on (None: leave it -- "philox" unless
    OPTPRICING_B200_RNG_MODE is set).  "numpy" replays the reference's own host draws, so seeded
    prices equal the reference's to 1e-12; "philox" is the throughput mode and meets the same
    contracts statistically."""
    reference_src = reference_src or os.environ.get("OPTPRICING_REFERENCE_SRC")
    if rng_mode is not None:
        from optpricing_b200.techniques.monte_carlo import set_default_rng_mode
        set_default_rng_mode(rng_mode)
    if "optpricing" not in sys.modules and reference_src:
        pkg_dir = os.path.join(reference_src, "optpricing")
        if not os.path.isdir(pkg_dir):
            raise FileNotFoundError(pkg_dir)
        shim = types.ModuleType("optpricing")
        shim.__path__ = [pkg_dir]
        shim.__version__ = "overlay"
        sys.modules["optpricing"] = shim
    for alias, target in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(target)
    # parents that were imported before install(): rebind what they already hold
    tech = sys.modules.get("optpricing.techniques")
    ours = importlib.import_module("optpricing_b200.techniques")
    if tech is not None:
        for name in _CLASSES:
            setattr(tech, name, getattr(ours, name))
        tech.monte_carlo = sys.modules["optpricing.techniques.monte_carlo"]
        tech.american_monte_carlo = sys.modules["optpricing.techniques.american_monte_carlo"]
    kern = sys.modules.get("optpricing.techniques.kernels")
    if kern is not None:
        for leaf in ("mc_kernels", "path_kernels", "american_mc_kernels"):
            mod = sys.modules[f"optpricing.techniques.kernels.{leaf}"]
            setattr(kern, leaf, mod)
            for name in getattr(mod, "__all__", None) or [n for n in vars(mod) if n.endswith(
                    ("_kernel", "_pricer"))]:
                if hasattr(kern, name):
                    setattr(kern, name, getattr(mod, name))
    top = sys.modules.get("optpricing")
    if top is not None:
        for name in _CLASSES:
            if hasattr(top, name):
                setattr(top, name, getattr(ours, name))
    # The one sampler the technique reaches through the model object
    # (monte_carlo.py:103-104): CEVModel.sample_terminal_spot -> device sampler.
    # The pure-Levy laws need no rebinding: _simulate_levy_terminal picks the
    # device sampler by model name and never calls sample_terminal_log_return.
    try:
        cev = importlib.import_module("optpricing.models.cev")
    except Exception:  # noqa: BLE001 - models not importable in this environment: nothing to rebind
        cev = None
    if cev is not None and hasattr(cev, "CEVModel"):
        from optpricing_b200.models import device_sample_terminal_spot
        cev.CEVModel.sample_terminal_spot = device_sample_terminal_spot


def installed() -> bool:
    mod = sys.modules.get("optpricing.techniques.monte_carlo")
    return mod is not None and mod.__name__ == "optpricing_b200.techniques.monte_carlo"
