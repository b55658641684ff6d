"""GPU-backed drop-ins for optpricing.techniques.kernels (kernels/__init__.py:11-61)."""
from .american_mc_kernels import DevicePathStore, longstaff_schwartz_pricer
from .mc_kernels import (bates_kernel, bsm_kernel, dupire_kernel, heston_kernel, kou_kernel,
                         merton_kernel, sabr_jump_kernel, sabr_kernel)
from .path_kernels import (bates_path_kernel, bsm_path_kernel, heston_path_kernel,
                           kou_path_kernel, merton_path_kernel, sabr_jump_path_kernel,
                           sabr_path_kernel)

__all__ = [
    "bsm_kernel", "heston_kernel", "merton_kernel", "bates_kernel", "sabr_kernel",
    "sabr_jump_kernel", "kou_kernel", "dupire_kernel", "bsm_path_kernel", "heston_path_kernel",
    "merton_path_kernel", "bates_path_kernel", "sabr_path_kernel", "sabr_jump_path_kernel",
    "kou_path_kernel", "longstaff_schwartz_pricer", "DevicePathStore",
]
