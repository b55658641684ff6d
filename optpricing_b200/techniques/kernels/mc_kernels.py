"""
Drop-in for the reference's terminal-value kernels
(/root/reference/src/optpricing/techniques/kernels/mc_kernels.py): same names,
same positional signatures, numpy in / numpy out -- but the step loop runs in
liboptmc.so's fed-draw kernel on the GPU.  Jump sizes, which the numba
originals pull from numba's hidden global RNG inside the loop, are drawn here
from numpy's global legacy stream (bit-identical for the same seed) and shipped
with the increments.
"""
from __future__ import annotations

from . import _fed


def bsm_kernel(n_paths, n_steps, log_s0, r, q, sigma, dt, dw):  # mc_kernels.py:18-34
    return _fed.run("bsm", {"sigma": sigma}, r, q, dt, log_s0, None, dw)


def heston_kernel(n_paths, n_steps, log_s0, v0, r, q, kappa, theta, rho, vol_of_vol, dt, dw1,
                  dw2):  # mc_kernels.py:43-79
    p = {"kappa": kappa, "theta": theta, "rho": rho, "vol_of_vol": vol_of_vol}
    return _fed.run("heston", p, r, q, dt, log_s0, v0, dw1, dw2)


def merton_kernel(n_paths, n_steps, log_s0, r, q, sigma, lambda_, mu_j, sigma_j, dt, dw,
                  jump_counts):  # mc_kernels.py:88-118
    p = {"sigma": sigma, "lambda": lambda_, "mu_j": mu_j, "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("merton", p, r, q, dt, log_s0, None, dw, None, jump_counts, jumps)


def bates_kernel(n_paths, n_steps, log_s0, v0, r, q, kappa, theta, rho, vol_of_vol, lambda_,
                 mu_j, sigma_j, dt, dw1, dw2, jump_counts):  # mc_kernels.py:127-177
    p = {"kappa": kappa, "theta": theta, "rho": rho, "vol_of_vol": vol_of_vol,
         "lambda": lambda_, "mu_j": mu_j, "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("bates", p, r, q, dt, log_s0, v0, dw1, dw2, jump_counts, jumps)


def sabr_kernel(n_paths, n_steps, s0, v0, r, q, alpha, beta, rho, dt, dw1,
                dw2):  # mc_kernels.py:186-218
    p = {"alpha": alpha, "beta": beta, "rho": rho}
    return _fed.run("sabr", p, r, q, dt, float(s0), v0, dw1, dw2)


def sabr_jump_kernel(n_paths, n_steps, s0, v0, r, q, alpha, beta, rho, lambda_, mu_j, sigma_j,
                     dt, dw1, dw2, jump_counts):  # mc_kernels.py:227-276
    p = {"alpha": alpha, "beta": beta, "rho": rho, "lambda": lambda_, "mu_j": mu_j,
         "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("sabr_jump", p, r, q, dt, float(s0), v0, dw1, dw2, jump_counts, jumps)


def kou_kernel(n_paths, n_steps, log_s0, r, q, sigma, lambda_, p_up, eta1, eta2, dt, dw,
               jump_counts):  # mc_kernels.py:285-329
    p = {"sigma": sigma, "lambda": lambda_, "p_up": p_up, "eta1": eta1, "eta2": eta2}
    jumps = _fed.jump_draws_kou(jump_counts, p_up, eta1, eta2)
    return _fed.run("kou", p, r, q, dt, log_s0, None, dw, None, jump_counts, jumps, sum_mode=1)


def dupire_kernel(n_paths, n_steps, log_s0, r, q, dt, dw, vol_surface_func, *, n_s=1024,
                  half_width=None):  # mc_kernels.py:340-359
    """
    The reference calls ``vol_surface_func(t_i, S)`` from inside its step loop;
    here the callable is sampled once on the host at the n_steps times and
    ``n_s`` log-spaced spots (``_engine.sample_local_vol``) and the device step
    interpolates linearly in log S.  Exact for surfaces that are piecewise
    linear in log S on that grid (in particular constant ones); otherwise the
    interpolation error is O(h^2 d2 sigma / d(log S)^2), h ~ 3e-3 by default.
    """
    eng = _fed.get_engine()
    lv = eng.local_vol_table(vol_surface_func, n_steps, dt, log_s0, r, q, n_s, half_width)
    return _fed.run("dupire", {}, r, q, dt, log_s0, None, dw, local_vol=lv)
