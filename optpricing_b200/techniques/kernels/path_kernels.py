"""
Drop-in for the reference's full-path kernels
(/root/reference/src/optpricing/techniques/kernels/path_kernels.py): each returns
the (n_paths, n_steps + 1) spot matrix, computed by liboptmc.so's fed-draw
kernel.  See mc_kernels.py in this package for the jump-size convention.
"""
from __future__ import annotations

from . import _fed


def bsm_path_kernel(n_paths, n_steps, log_s0, r, q, sigma, dt, dw):  # path_kernels.py:17-35
    return _fed.run("bsm", {"sigma": sigma}, r, q, dt, log_s0, None, dw, paths=True)


def heston_path_kernel(n_paths, n_steps, log_s0, v0, r, q, kappa, theta, rho, vol_of_vol, dt,
                       dw1, dw2):  # path_kernels.py:43-77
    p = {"kappa": kappa, "theta": theta, "rho": rho, "vol_of_vol": vol_of_vol}
    return _fed.run("heston", p, r, q, dt, log_s0, v0, dw1, dw2, paths=True)


def merton_path_kernel(n_paths, n_steps, log_s0, r, q, sigma, lambda_, mu_j, sigma_j, dt, dw,
                       jump_counts):  # path_kernels.py:85-121
    p = {"sigma": sigma, "lambda": lambda_, "mu_j": mu_j, "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("merton", p, r, q, dt, log_s0, None, dw, None, jump_counts, jumps,
                    paths=True, sum_mode=1)


def bates_path_kernel(n_paths, n_steps, log_s0, v0, r, q, kappa, theta, rho, vol_of_vol,
                      lambda_, mu_j, sigma_j, dt, dw1, dw2,
                      jump_counts):  # path_kernels.py:130-180
    p = {"kappa": kappa, "theta": theta, "rho": rho, "vol_of_vol": vol_of_vol,
         "lambda": lambda_, "mu_j": mu_j, "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("bates", p, r, q, dt, log_s0, v0, dw1, dw2, jump_counts, jumps, paths=True,
                    sum_mode=1)


def sabr_path_kernel(n_paths, n_steps, s0, v0, r, q, alpha, beta, rho, dt, dw1,
                     dw2):  # path_kernels.py:189-217
    p = {"alpha": alpha, "beta": beta, "rho": rho}
    return _fed.run("sabr", p, r, q, dt, float(s0), v0, dw1, dw2, paths=True)


def sabr_jump_path_kernel(n_paths, n_steps, s0, v0, r, q, alpha, beta, rho, lambda_, mu_j,
                          sigma_j, dt, dw1, dw2, jump_counts):  # path_kernels.py:225-274
    p = {"alpha": alpha, "beta": beta, "rho": rho, "lambda": lambda_, "mu_j": mu_j,
         "sigma_j": sigma_j}
    jumps = _fed.jump_draws_normal(jump_counts, mu_j, sigma_j)
    return _fed.run("sabr_jump", p, r, q, dt, float(s0), v0, dw1, dw2, jump_counts, jumps,
                    paths=True)


def kou_path_kernel(n_paths, n_steps, log_s0, r, q, sigma, lambda_, p_up, eta1, eta2, dt, dw,
                    jump_counts):  # path_kernels.py:277-324
    p = {"sigma": sigma, "lambda": lambda_, "p_up": p_up, "eta1": eta1, "eta2": eta2}
    jumps = _fed.jump_draws_kou(jump_counts, p_up, eta1, eta2)
    return _fed.run("kou", p, r, q, dt, log_s0, None, dw, None, jump_counts, jumps, paths=True,
                    sum_mode=1)
