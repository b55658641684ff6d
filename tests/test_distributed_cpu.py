"""
world_size-2 gloo tests (CPU) of the multi-rank host logic: draw partition,
all-reduce of the European moments, and the per-date all-reduce placement of the
Longstaff-Schwartz loop.  The CUDA library is replaced by numpy test doubles that
follow the C ABI's contract (include/optmc.h); the arithmetic of the doubles
comes from the oracle.  The real kernels are exercised by the GPU tests and by
tests/dist_gpu_check.py under torchrun.
"""
import ctypes
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import optpricing_b200 as ob
from optpricing_b200 import _engine
from optpricing_b200.techniques import monte_carlo as mc_mod
from oracle import reference_mc as orc

N_PATHS, N_STEPS = 4000, 6
HESTON = ob.HestonModel().params


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _job_terminals(seed):
    """Terminal spots of the whole job, draw-major (originals, then partners)."""
    rng = np.random.default_rng(seed)
    _, ST = orc.price_european("heston", HESTON, 100.0, 100.0, 1.0, 0.05, 0.01, True, N_PATHS,
                               N_STEPS, True, rng, return_terminal=True)
    half = N_PATHS // 2
    return ST[:half], ST[half:]


class FakeEuropeanEngine:
    """Stands in for _engine.Engine: moments of this rank's draw range."""
    device = torch.device("cpu")

    def __init__(self):
        self._buf = {}

    def buffer(self, name, n, dtype=None):
        return self._buf.setdefault(name, torch.zeros(n, dtype=torch.float64))

    def read(self, t, n):
        return t[:n].cpu().numpy()

    def make_sim(self, s0, maturity, n_draws, n_steps, antithetic, seed, correlation,
                 draw_offset=0, n_local=None, precision=0):
        return dict(seed=seed, off=draw_offset, cnt=n_local, n_draws=n_draws)

    def _moments(self, sim, K, call):
        a, b = _job_terminals(sim["seed"] % 1000)
        sl = slice(sim["off"], sim["off"] + sim["cnt"])
        pay = lambda s: np.maximum(s - K, 0) if call else np.maximum(K - s, 0)  # noqa: E731
        p = 0.5 * (pay(a[sl]) + pay(b[sl]))
        c = 0.5 * (a[sl] + b[sl])
        return np.array([p.size, p.sum(), (p * p).sum(), c.sum(), (c * c).sum(), (p * c).sum()])

    def european(self, mdl, sim, strikes, is_call, terminal=None):
        return self._moments(sim, strikes[0], bool(is_call[0]))[None, :]

    def european_async(self, mdl, sim, strikes, is_call, moments, terminal=None, aux=None):
        for j, (k, c) in enumerate(zip(strikes, is_call)):          # strike sweep: 6 per contract
            moments[6 * j:6 * j + 6] = torch.from_numpy(self._moments(sim, float(k), bool(c)))


def _european_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _engine.get_engine = lambda device=None: FakeEuropeanEngine()
        _engine.make_model = lambda *a, **k: None
        t = ob.MonteCarloTechnique(n_paths=N_PATHS, n_steps=N_STEPS, seed=11)
        opt = ob.Option(100.0, 1.0, ob.OptionType.CALL)
        price = t.price(opt, ob.Stock(100.0, dividend=0.01), ob.HestonModel(), ob.Rate(0.05)).price
        out[rank] = (price, t.last_stats["n_samples"], t.last_stats["seed"])
    finally:
        dist.destroy_process_group()


def _sweep_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _engine.get_engine = lambda device=None: FakeEuropeanEngine()
        _engine.make_model = lambda *a, **k: None
        t = ob.MonteCarloTechnique(n_paths=N_PATHS, n_steps=N_STEPS, seed=11)
        opts = [ob.Option(k, 1.0, ob.OptionType.CALL if k < 105 else ob.OptionType.PUT)
                for k in (90.0, 100.0, 110.0)]
        prices = t.price_many(opts, ob.Stock(100.0, dividend=0.01), ob.HestonModel(), ob.Rate(0.05))
        out[rank] = (list(map(float, prices)), t.last_stats_many[1]["n_samples"],
                     t.last_stats_many[0]["seed"])
    finally:
        dist.destroy_process_group()


def test_strike_sweep_one_allreduce_two_ranks():
    """price_many under two ranks: every rank sweeps its own draws, ONE all-reduce of all
    moments, and both ranks hold the whole-job prices."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sweep_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    (p0, n0, s0), (p1, n1, s1) = res[0], res[1]
    assert p0 == p1 and s0 == s1 and n0 == n1 == N_PATHS // 2
    a, b = _job_terminals(s0 % 1000)
    disc = math.exp(-0.05)
    want = [disc * np.mean(0.5 * (np.maximum(a - 90, 0) + np.maximum(b - 90, 0))),
            disc * np.mean(0.5 * (np.maximum(a - 100, 0) + np.maximum(b - 100, 0))),
            disc * np.mean(0.5 * (np.maximum(110 - a, 0) + np.maximum(110 - b, 0)))]
    assert p0 == pytest.approx(want, rel=1e-13)


def test_european_moments_allreduce_two_ranks():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_european_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    (p0, n0, s0), (p1, n1, s1) = res[0], res[1]
    assert p0 == p1 and s0 == s1 and n0 == n1 == N_PATHS // 2
    a, b = _job_terminals(s0 % 1000)
    want = math.exp(-0.05) * np.mean(0.5 * (np.maximum(a - 100, 0) + np.maximum(b - 100, 0)))
    assert p0 == pytest.approx(want, rel=1e-13)


# ---- Longstaff-Schwartz: per-date all-reduce of the power sums ---------------
class FakeLsmcLib:
    """numpy double of optmc_lsmc_step_async (include/optmc.h), on host pointers."""

    def __init__(self):
        self.calls = []

    @staticmethod
    def _arr(ptr, n):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), (n,))

    def optmc_last_error(self, ctx):
        return b""

    def optmc_lsmc_step_async(self, ctx, store, ld, n, n_steps, date, K, is_call, disc, deg, cash,
                              gram, stream):
        self.calls.append(date)
        S = self._arr(store, ld * n_steps).reshape(n_steps, ld)[:, :n]
        cf = self._arr(cash, n)
        G = self._arr(gram, 24 * n_steps).reshape(n_steps, 24)
        iv = lambda s: (s - K) if is_call else (K - s)  # noqa: E731
        if date == n_steps - 1:
            y = np.maximum(iv(S[date]), 0.0)
        else:
            y = cf.copy()
            g = G[date + 1]
            m = deg + 1
            if g[0] >= m:
                A = np.array([[g[a + b] for b in range(m)] for a in range(m)])
                rhs = g[16:16 + m]
                try:
                    beta = np.linalg.solve(A, rhs)
                    x = S[date] / K - 1.0
                    cont = sum(beta[k] * x**k for k in range(m))
                    ex = (iv(S[date]) > 0) & (iv(S[date]) > cont)
                    y = np.where(ex, iv(S[date]), y)
                except np.linalg.LinAlgError:
                    pass
        y = y * disc
        G[date] = 0.0
        if date >= 1:
            cf[:] = y
            itm = iv(S[date - 1]) > 0
            x = S[date - 1][itm] / K - 1.0
            for k in range(2 * deg + 1):
                G[date, k] = np.sum(x**k)
            for k in range(deg + 1):
                G[date, 16 + k] = np.sum(y[itm] * x**k)
        else:
            G[0, :3] = [n, y.sum(), (y * y).sum()]
        return 0


def _lsmc_paths():
    return orc.simulate_paths("bsm", {"sigma": 0.2}, 100.0, 0.05, 0.01, 1.0, 3000, 8, True,
                              np.random.default_rng(5))


def _lsmc_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        paths = _lsmc_paths()
        lo, cnt = mc_mod.shard_draws(paths.shape[0], rank, world)
        mine = paths[lo:lo + cnt]
        n, n_steps = mine.shape[0], mine.shape[1] - 1
        eng = _engine.Engine.__new__(_engine.Engine)     # no CUDA: wire the double in by hand
        eng.torch, eng.lib, eng.ctx, eng._buffers = torch, FakeLsmcLib(), None, {}
        eng.device = torch.device("cpu")
        eng.stream = lambda: None
        ld = (n + 1) // 2 * 2
        store = torch.zeros(ld * n_steps, dtype=torch.float64)
        store.view(n_steps, ld)[:, :n] = torch.from_numpy(np.ascontiguousarray(mine[:, 1:].T))
        res = eng.lsmc_distributed(store, ld, n, n_steps, 110.0, False, 0.05, 1.0 / n_steps, 2,
                                   dist.all_reduce)
        out[rank] = (float(res[1] / res[0]), list(eng.lib.calls))
    finally:
        dist.destroy_process_group()


def test_lsmc_gram_allreduce_two_ranks():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_lsmc_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    (p0, calls0), (p1, calls1) = res[0], res[1]
    assert p0 == p1
    assert calls0 == calls1 == list(range(7, -1, -1))
    paths = _lsmc_paths()
    want = orc.longstaff_schwartz(paths, 110.0, 1.0 / 8, 0.05, False, 2)
    assert p0 == pytest.approx(want, rel=1e-9)
