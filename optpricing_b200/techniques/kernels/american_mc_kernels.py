"""
Drop-in for the reference's Longstaff-Schwartz pricer
(/root/reference/src/optpricing/techniques/kernels/american_mc_kernels.py:36-84).

``stock_paths`` may be the reference's host matrix (n_paths, n_steps + 1), which
is copied to the GPU and transposed into the step-major store, or a
``DevicePathStore`` produced by AmericanMonteCarloTechnique, in which case the
paths never leave HBM.
"""
from __future__ import annotations

import numpy as np

from optpricing_b200._engine import get_engine, serialized


class DevicePathStore:
    """Handle to a step-major path store in HBM (rows = dates 1..n_steps)."""

    def __init__(self, engine, store, ld, n_paths, n_steps, s0):
        self.engine, self.store, self.ld = engine, store, ld
        self.n_paths, self.n_steps, self.s0 = n_paths, n_steps, s0
        self.shape = (n_paths, n_steps + 1)

    def __array__(self, dtype=None, copy=None):
        """Materialise the reference-layout matrix on the host (diagnostics only)."""
        rows = self.store[: self.ld * self.n_steps].view(self.n_steps, self.ld)[:, : self.n_paths]
        out = np.empty(self.shape)
        out[:, 0] = self.s0
        out[:, 1:] = rows.t().cpu().numpy()
        return out if dtype is None else out.astype(dtype)


@serialized
def longstaff_schwartz_pricer(stock_paths, K, dt, r, is_call, degree, all_reduce=None,
                              peers=False):
    """
    Returns the price as a float.  With several ranks the paths given here are
    this rank's shard and the per-date power sums are exchanged either inside the
    kernel over NVLink peer memory (``peers=True``, after
    ``engine.lsmc_peer_attach()``) or by ``all_reduce`` (sums a device tensor
    across ranks in place, e.g. NCCL) between per-date launches.
    """
    if isinstance(stock_paths, DevicePathStore):
        h = stock_paths
        eng = h.engine
    else:
        eng = get_engine()
        torch = eng.torch
        host = np.ascontiguousarray(stock_paths, dtype=np.float64)
        n_paths, n_steps = host.shape[0], host.shape[1] - 1
        dev = torch.from_numpy(host).to(eng.device)
        if n_steps < 1:
            raise ValueError("stock_paths needs at least two columns")
        store, ld = eng.lsmc_store(n_paths, n_steps)
        eng.lsmc_load(dev, n_paths, n_steps, store, ld)
        h = DevicePathStore(eng, store, ld, n_paths, n_steps, float(host[0, 0]))
    if peers or all_reduce is None:
        n, s, _ = eng.lsmc(h.store, h.ld, h.n_paths, h.n_steps, K, bool(is_call), r, dt, degree,
                           peers=peers)
    else:
        n, s, _ = eng.lsmc_distributed(h.store, h.ld, h.n_paths, h.n_steps, K, bool(is_call), r,
                                       dt, degree, all_reduce)
    return float(s / n)
