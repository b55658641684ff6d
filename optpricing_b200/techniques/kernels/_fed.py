"""Shared plumbing of the fed-draw kernel wrappers (mc_kernels / path_kernels)."""
from __future__ import annotations

import numpy as np

from optpricing_b200._engine import get_engine, make_model, serialized


def jump_draws_normal(jump_counts, mu_j, sigma_j):
    """
    The jump sizes a reference kernel pass consumes, drawn the way it draws
    them -- one batch per step that has jumps (mc_kernels.py:108-111) -- from
    numpy's legacy global stream, which is bit-identical to the numba global
    stream the reference uses (tests/golden/make_golden.py asserts this).
    """
    per_step = np.asarray(jump_counts).sum(axis=0)
    chunks = [np.random.normal(mu_j, sigma_j, int(n)) for n in per_step if n > 0]
    return np.concatenate(chunks) if chunks else np.zeros(0)


def jump_draws_kou(jump_counts, p_up, eta1, eta2):
    """mc_kernels.py:309-317: uniforms, up sizes, down sizes, in that order per step."""
    per_step = np.asarray(jump_counts).sum(axis=0)
    chunks = []
    for n in per_step:
        n = int(n)
        if n > 0:
            u = np.random.random(n)
            up = np.random.exponential(1.0 / eta1, n)
            dn = -np.random.exponential(1.0 / eta2, n)
            chunks.append(np.where(u < p_up, up, dn))
    return np.concatenate(chunks) if chunks else np.zeros(0)


@serialized
def run(kind, params, r, q, dt, x0, v0, dw1, dw2=None, jump_counts=None, jumps=None, *,
        paths=False, sum_mode=0, local_vol=None):
    eng = get_engine()
    model = make_model(kind, params, r, q, v0=v0, local_vol=local_vol)
    term, pth = eng.fed(kind, model, x0, dt, dw1, dw2, jump_counts, jumps, sign=1.0,
                        path_sum_mode=sum_mode, want_terminal=not paths, want_paths=paths)
    return pth if paths else term
