#!/usr/bin/env python
"""GPU experiment: HBM throughput of the fed-draw kernels (the parity / rng_mode="numpy" path)
with the increments already resident on the device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, math
import torch
from optpricing_b200 import _engine

eng = _engine.get_engine()
n, m = int(os.environ.get("N_PATHS", 1 << 20)), int(os.environ.get("N_STEPS", 252))
dt = 1.0 / m
g = torch.Generator(device=eng.device).manual_seed(1)
dw1 = torch.randn((n, m), dtype=torch.float64, device=eng.device, generator=g) * math.sqrt(dt)
dw2 = torch.randn((n, m), dtype=torch.float64, device=eng.device, generator=g) * math.sqrt(dt)
term = torch.empty(n, dtype=torch.float64, device=eng.device)
paths = torch.empty((n, m + 1), dtype=torch.float64, device=eng.device)
ws = torch.empty(1, dtype=torch.int64, device=eng.device)
ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
lam_dt = 0.5 * dt
counts = torch.poisson(torch.full((n, m), lam_dt, dtype=torch.float64, device=eng.device), generator=g).to(torch.int64)
n_jumps = int(counts.sum().item())
jumps = torch.randn(max(n_jumps, 1), dtype=torch.float64, device=eng.device, generator=g) * 0.15 - 0.1
HES = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5}
JMP = {"lambda": 0.5, "mu_j": -0.1, "sigma_j": 0.15}
for kind, params, two in (("bsm", {"sigma": 0.2}, False), ("heston", HES, True),
                          ("merton", {"sigma": 0.2, **JMP}, False), ("bates", {**HES, **JMP}, True)):
    jump = kind in ("merton", "bates")
    if jump:
        ws = torch.empty(max(int(eng.lib.optmc_fed_workspace_bytes(_engine.KIND_CODES[kind], n, m)) // 8, 1),
                         dtype=torch.int64, device=eng.device)
    mdl = _engine.make_model(kind, params, 0.05, 0.0, v0=0.04 if two else None)
    for want_paths in (False, True):
        best = 1e9
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng._check(eng.lib.optmc_fed_async(
                eng.ctx, ctypes.byref(mdl), math.log(100.0), dt, n, m, ptr(dw1), ptr(dw2) if two else None,
                ptr(counts) if jump else None, ptr(jumps) if jump else None, n_jumps if jump else 0,
                1.0, 0, ptr(term) if not want_paths else None,
                ptr(paths) if want_paths else None, ptr(ws), ws.numel() * 8, eng.stream()),
                "optmc_fed_async")
            b.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, a.elapsed_time(b))
        bytes_ = n * m * 8 * ((2 if two else 1) + (1 if want_paths else 0) + (1 if jump else 0))
        print(f"k_fed<{kind}> {n} x {m} {'paths' if want_paths else 'terminal'}: {best:.3f} ms = "
              f"{bytes_ / best / 1e6:.0f} GB/s ({bytes_ / 2**30:.1f} GiB moved)", flush=True)
