"""
ctypes binding of liboptmc.so (include/optmc.h) plus the per-process engine
object the techniques share.

PyTorch is plumbing here: it owns device buffers (``torch.empty(..., device=)``)
and the stream; every kernel is in liboptmc.so.  There is no CPU fallback: if
the library is missing or no B200 is visible, constructing the engine raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OPTMC_LIB") or os.path.join(_HERE, "liboptmc.so")

KIND_CODES = {"bsm": 0, "heston": 1, "merton": 2, "bates": 3, "sabr": 4, "sabr_jump": 5, "kou": 6,
              "dupire": 7}
TWO_FACTOR = ("heston", "bates", "sabr", "sabr_jump")
JUMP_KINDS = ("merton", "bates", "sabr_jump", "kou")
SABR_KINDS = ("sabr", "sabr_jump")
CORR_REFERENCE_EUROPEAN = 0
CORR_INDEPENDENT_DRAWS = 1
NMOM = 6
LSMC_GRAM = 24
MAX_PEERS = 8
IPC_HANDLE_BYTES = 64
LEVY_CODES = {"vg": 0, "nig": 1, "cgmy": 2, "hyperbolic": 3, "cev": 4, "lognormal": 5}
LEVY_NPARAM = 12

c_void_p, c_int, c_i32, c_i64, c_u32, c_u64, c_double = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
    ctypes.c_uint64, ctypes.c_double)
_dp = ctypes.POINTER(c_double)
_ip = ctypes.POINTER(c_i32)


class OptmcModel(ctypes.Structure):
    """struct optmc_model (include/optmc.h)."""
    _fields_ = [("kind", c_i32), ("reserved", c_i32), ("r", c_double), ("q", c_double),
                ("sigma", c_double), ("v0", c_double), ("kappa", c_double), ("theta", c_double),
                ("rho", c_double), ("vol_of_vol", c_double), ("alpha", c_double),
                ("beta", c_double), ("lambda_", c_double), ("mu_j", c_double),
                ("sigma_j", c_double), ("p_up", c_double), ("eta1", c_double),
                ("eta2", c_double), ("lv_table_dev", c_void_p), ("lv_n_s", c_i32),
                ("lv_n_t", c_i32), ("lv_log_s_lo", c_double), ("lv_inv_h", c_double)]


class OptmcSim(ctypes.Structure):
    """struct optmc_sim (include/optmc.h)."""
    _fields_ = [("s0", c_double), ("maturity", c_double), ("n_draws", c_i64),
                ("draw_offset", c_i64), ("n_local", c_i64), ("n_steps", c_i32),
                ("antithetic", c_i32), ("seed", c_u64), ("correlation", c_i32),
                ("precision", c_i32)]


class OptmcLevy(ctypes.Structure):
    """struct optmc_levy (include/optmc.h)."""
    _fields_ = [("kind", c_i32), ("reserved", c_i32), ("p", c_double * LEVY_NPARAM)]


# every symbol include/optmc.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("optmc_version", c_int, []),
    ("optmc_ctx_create", c_int, [c_int, ctypes.POINTER(c_void_p)]),
    ("optmc_ctx_destroy", None, [c_void_p]),
    ("optmc_last_error", ctypes.c_char_p, [c_void_p]),
    ("optmc_european_async", c_int, [c_void_p, ctypes.POINTER(OptmcModel),
                                     ctypes.POINTER(OptmcSim), c_i32, _dp, _ip, c_void_p,
                                     c_void_p, c_void_p]),
    ("optmc_european", c_int, [c_void_p, ctypes.POINTER(OptmcModel), ctypes.POINTER(OptmcSim),
                               c_i32, _dp, _ip, _dp, c_void_p, c_void_p]),
    ("optmc_fed_workspace_bytes", c_i64, [c_i32, c_i64, c_i32]),
    ("optmc_fed_async", c_int, [c_void_p, ctypes.POINTER(OptmcModel), c_double, c_double, c_i64,
                                c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_double,
                                c_i32, c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    ("optmc_european_aux_async", c_int, [c_void_p, ctypes.POINTER(OptmcModel),
                                         ctypes.POINTER(OptmcSim), c_i32, _dp, _ip, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    ("optmc_european_scen_async", c_int, [c_void_p, ctypes.POINTER(OptmcModel),
                                          ctypes.POINTER(OptmcSim), c_i32, _dp, c_double, c_i32,
                                          c_void_p, c_void_p]),
    ("optmc_peer_allreduce_async", c_int, [c_void_p, c_void_p, c_i32, c_void_p]),
    ("optmc_european_allreduce", c_int, [c_void_p, ctypes.POINTER(OptmcModel),
                                         ctypes.POINTER(OptmcSim), c_i32, _dp, _ip, _dp, c_void_p,
                                         c_void_p]),
    ("optmc_read_doubles", c_int, [c_void_p, c_void_p, c_i32, _dp, c_void_p]),
    ("optmc_sigma_sweep_async", c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, _dp, _dp, _dp,
                                        _ip, c_void_p, c_void_p]),
    ("optmc_payoff_sweep_async", c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, _dp, _ip,
                                         c_void_p, c_void_p]),
    ("optmc_levy_gig_setup", c_int, [ctypes.POINTER(OptmcLevy)]),
    ("optmc_levy_terminal_async", c_int, [c_void_p, ctypes.POINTER(OptmcLevy), c_double, c_double,
                                          c_u64, c_i64, c_i64, c_i32, _dp, _ip, c_void_p,
                                          c_void_p, c_void_p]),
    ("optmc_lsmc_paths_async", c_int, [c_void_p, ctypes.POINTER(OptmcModel),
                                       ctypes.POINTER(OptmcSim), c_void_p, c_i64, c_void_p]),
    ("optmc_lsmc_load_async", c_int, [c_void_p, c_void_p, c_i64, c_i32, c_void_p, c_i64,
                                      c_void_p]),
    ("optmc_lsmc_step_async", c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i32, c_i32, c_double,
                                      c_i32, c_double, c_i32, c_void_p, c_void_p, c_void_p]),
    ("optmc_lsmc", c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i32, c_double, c_i32, c_double,
                           c_double, c_i32, c_void_p, c_void_p, _dp, c_void_p]),
    ("optmc_lsmc_peer_export", c_int, [c_void_p, c_void_p]),
    ("optmc_lsmc_peer_attach", c_int, [c_void_p, c_i32, c_i32, c_void_p]),
    ("optmc_philox_raw", c_int, [c_void_p, c_u64, c_i64, c_u32, c_u32, c_void_p, c_void_p]),
    ("optmc_normals", c_int, [c_void_p, c_u64, c_i64, c_i32, c_void_p, c_void_p]),
    ("optmc_fp64_peak", c_int, [c_void_p, c_i64, _dp, _dp]),
    ("optmc_set_int", c_int, [c_void_p, ctypes.c_char_p, c_i64]),
    ("optmc_set_ptr", c_int, [c_void_p, ctypes.c_char_p, c_void_p]),
    ("optmc_launch_count", c_i64, [c_void_p]),
]


class EngineError(RuntimeError):
    pass


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen liboptmc.so and bind every declared symbol.  Needs no GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} is missing: build it with `python -m optpricing_b200.build` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SIGNATURES:
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


class LocalVolTable:
    """
    A Dupire local-vol surface sampled for one simulation: ``table`` is a device
    tensor [n_t, n_s] with table[i, j] = vol_surface(i * dt, exp(lo + j * h)).
    The model struct keeps a raw pointer into it, so hold on to this object for
    as long as the struct is in use.
    """

    def __init__(self, table, lo: float, h: float):
        self.table, self.lo, self.h = table, float(lo), float(h)
        self.n_t, self.n_s = int(table.shape[0]), int(table.shape[1])


def sample_local_vol(vol_surface, n_steps: int, dt: float, log_s0: float, r: float, q: float,
                     n_s: int = 1024, half_width: float | None = None):
    """
    Host side of the Dupire kernel: evaluate ``vol_surface(t, S)`` -- the callable
    the reference calls from inside its step loop (mc_kernels.py:351-357) -- at
    the n_steps times the kernel uses and n_s spots equally spaced in log S.
    Returns (numpy table [n_steps, n_s], lo, h).  The grid spans
    log S0 +- half_width (default: 10 standard deviations of the largest
    at-the-money vol plus the drift, at least 1.5); beyond it the surface is flat.
    """
    import math

    T = n_steps * dt
    times = [i * dt for i in range(n_steps)]
    if half_width is None:
        s0 = math.exp(log_s0)
        atm = max(abs(float(np.max(np.asarray(vol_surface(t, np.array([s0])), dtype=float))))
                  for t in times[:: max(1, n_steps // 16)])
        half_width = max(1.5, 10.0 * atm * math.sqrt(T) + abs(r - q) * T)
    lo, hi = log_s0 - half_width, log_s0 + half_width
    grid = np.linspace(lo, hi, int(n_s))
    spots = np.exp(grid)
    table = np.empty((n_steps, int(n_s)))
    for i, t in enumerate(times):
        table[i] = np.broadcast_to(np.asarray(vol_surface(t, spots), dtype=float), spots.shape)
    if not np.all(np.isfinite(table)):
        raise ValueError("vol_surface returned non-finite local volatilities")
    return table, lo, (hi - lo) / (int(n_s) - 1)


def make_model(kind: str, params: dict, r: float, q: float, v0=None,
               local_vol: LocalVolTable | None = None) -> OptmcModel:
    """
    Flatten the reference's ``kernel_params`` (monte_carlo.py:128-197) into
    struct optmc_model.  ``v0`` follows the reference: kwarg first, then
    params["v0"] (Heston/Bates) or params["alpha"] (SABR).
    """
    m = OptmcModel()
    m.kind = KIND_CODES[kind]
    m.r, m.q = float(r), float(q)
    p = params
    if kind == "dupire":
        if local_vol is None:
            raise ValueError("the Dupire kernel needs a sampled surface (sample_local_vol)")
        m.lv_table_dev = local_vol.table.data_ptr()
        m.lv_n_s, m.lv_n_t = local_vol.n_s, local_vol.n_t
        m.lv_log_s_lo, m.lv_inv_h = local_vol.lo, 1.0 / local_vol.h
        return m
    if kind in ("bsm", "merton", "kou"):
        m.sigma = float(p["sigma"])
    if kind in ("heston", "bates"):
        m.kappa, m.theta, m.rho, m.vol_of_vol = (float(p["kappa"]), float(p["theta"]),
                                                 float(p["rho"]), float(p["vol_of_vol"]))
        m.v0 = float(v0 if v0 is not None else p.get("v0"))
    if kind in SABR_KINDS:
        m.alpha, m.beta, m.rho = float(p["alpha"]), float(p["beta"]), float(p["rho"])
        m.v0 = float(v0 if v0 is not None else p.get("alpha"))
    if kind in ("merton", "bates", "sabr_jump"):
        m.lambda_, m.mu_j, m.sigma_j = float(p["lambda"]), float(p["mu_j"]), float(p["sigma_j"])
    if kind == "kou":
        m.lambda_, m.p_up, m.eta1, m.eta2 = (float(p["lambda"]), float(p["p_up"]),
                                             float(p["eta1"]), float(p["eta2"]))
    return m


def make_levy(kind: str, params: dict, r: float, T: float, S0: float) -> OptmcLevy:
    """
    Fold a single-step model's parameters into struct optmc_levy, in the
    reference's own parameterisation of each sampler:
    vg.py:164-168, nig.py:163-169, cgmy.py:174-178, hyperbolic.py:143-152
    (scipy genhyperbolic._rvs: GIG(p, sqrt(a^2-b^2)) scaled by 1/sqrt(a^2-b^2)),
    cev.py:80-94.
    """
    import math

    L = OptmcLevy()
    p = params
    vals: list[float]
    if kind == "vg":
        vals = [T / p["nu"], p["nu"], p["theta"], p["sigma"]]
    elif kind == "nig":
        vals = [p["delta"] * T / math.sqrt(p["alpha"] ** 2 - p["beta"] ** 2), p["beta"]]
    elif kind == "cgmy":
        if not math.isclose(p["Y"], 1.0, rel_tol=1e-5, abs_tol=1e-8):        # cgmy.py:168-171
            raise NotImplementedError("Monte Carlo sampling for CGMY is only implemented for Y=1.")
        vals = [p["C"] * T, 1.0 / p["M"], 1.0 / p["G"]]
    elif kind == "hyperbolic":
        t1 = p["alpha"] ** 2 - p["beta"] ** 2
        vals = [math.sqrt(t1), 1.0 / math.sqrt(t1), p["beta"], p["mu"] * T, p["delta"] * T, 0.0]
    elif kind == "cev":
        sigma, gamma = p["sigma"], p["gamma"]
        if abs(1.0 - gamma) < 1e-6:                                           # cev.py:81-85
            kind = "lognormal"
            vals = [(r - 0.5 * sigma ** 2) * T, sigma * math.sqrt(T)]
        else:
            df = 2 + 2 / (1 - gamma)
            k = 2 * r / (sigma ** 2 * (1 - gamma) * (math.exp(r * (1 - gamma) * T) - 1))
            nc = k * S0 ** (2 * (1 - gamma)) * math.exp(-r * (1 - gamma) * T)
            vals = [df, nc, k, 1 / (2 * (1 - gamma))]
    else:
        raise KeyError(kind)
    L.kind = LEVY_CODES[kind]
    for i, v in enumerate(vals):
        L.p[i] = float(v)
    if kind == "hyperbolic" and load_library().optmc_levy_gig_setup(ctypes.byref(L)) != 0:
        raise EngineError("optmc_levy_gig_setup: parameters outside the sampler's range")
    return L


class Engine:
    """One liboptmc context on one CUDA device, plus torch-owned device buffers."""

    def __init__(self, device: int | None = None):
        import torch

        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device visible: the optpricing_b200 engine needs a B200 "
                              "(there is no CPU fallback)")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        ctx = c_void_p()
        rc = self.lib.optmc_ctx_create(self.device_index, ctypes.byref(ctx))
        if rc != 0:
            raise EngineError("optmc_ctx_create: " + self.lib.optmc_last_error(None).decode())
        self.ctx = ctx
        self._buffers = {}
        self.n_simulations = 0      # European simulations launched (bench.py counts path-steps with it)

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what}: {self.lib.optmc_last_error(self.ctx).decode()}")

    def stream(self):
        return c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def buffer(self, name, n, dtype=None):
        """A cached device buffer of at least n elements (grown, never shrunk)."""
        torch = self.torch
        dtype = dtype or torch.float64
        t = self._buffers.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(int(n), 1), dtype=dtype, device=self.device)
            self._buffers[name] = t
        return t

    def set_int(self, key: str, value: int):
        self._check(self.lib.optmc_set_int(self.ctx, key.encode(), int(value)), "optmc_set_int")

    def set_ptr(self, key: str, tensor_or_none):
        ptr = c_void_p(tensor_or_none.data_ptr()) if tensor_or_none is not None else None
        self._check(self.lib.optmc_set_ptr(self.ctx, key.encode(), ptr), "optmc_set_ptr")

    def launch_count(self) -> int:
        return int(self.lib.optmc_launch_count(self.ctx))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.optmc_ctx_destroy(self.ctx)
            self.ctx = None

    # -- European, device RNG ----------------------------------------------
    def make_sim(self, s0, maturity, n_draws, n_steps, antithetic, seed, correlation,
                 draw_offset=0, n_local=None, precision=0) -> OptmcSim:
        s = OptmcSim()
        s.s0, s.maturity = float(s0), float(maturity)
        s.n_draws = int(n_draws)
        s.draw_offset = int(draw_offset)
        s.n_local = int(n_draws - draw_offset if n_local is None else n_local)
        s.n_steps = int(n_steps)
        s.antithetic = 1 if antithetic else 0
        s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        s.correlation = int(correlation)
        s.precision = int(precision)          # 0 = fp64, 1 = fp32 (enum optmc_precision)
        return s

    @staticmethod
    def _contracts(strikes, is_call):
        if len(strikes) == 1 and len(is_call) == 1:          # one contract: no numpy round trip
            k = np.array([strikes[0]], dtype=np.float64)
            c = np.array([1 if is_call[0] else 0], dtype=np.int32)
            return k, c, k.ctypes.data_as(_dp), c.ctypes.data_as(_ip)
        k = np.ascontiguousarray(np.atleast_1d(strikes), dtype=np.float64)
        c = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(is_call), k.shape), dtype=np.int32)
        return k, c, k.ctypes.data_as(_dp), c.ctypes.data_as(_ip)

    def european(self, model: OptmcModel, sim: OptmcSim, strikes, is_call, terminal=None):
        """optmc_european: host moments array (n_contracts, 6) after one sync."""
        self.n_simulations += 1
        k, c, kp, cp = self._contracts(strikes, is_call)
        out = np.empty((k.size, NMOM))
        tptr = c_void_p(terminal.data_ptr()) if terminal is not None else None
        if k.size > 1 and terminal is None:
            terminal = self.buffer("terminal", sim.n_local * (2 if sim.antithetic else 1))
            tptr = c_void_p(terminal.data_ptr())
        self._check(self.lib.optmc_european(self.ctx, ctypes.byref(model), ctypes.byref(sim),
                                            k.size, kp, cp, out.ctypes.data_as(_dp), tptr,
                                            self.stream()), "optmc_european")
        return out

    def european_allreduce(self, model: OptmcModel, sim: OptmcSim, strikes, is_call):
        """optmc_european_allreduce: this rank's shard, the all-reduce over the attached ranks and
        the read in one call -> host moments (n_contracts, 6), global and identical on all ranks."""
        self.n_simulations += 1
        k, c, kp, cp = self._contracts(strikes, is_call)
        out = np.empty((k.size, NMOM))
        self._check(self.lib.optmc_european_allreduce(self.ctx, ctypes.byref(model),
                                                      ctypes.byref(sim), k.size, kp, cp,
                                                      out.ctypes.data_as(_dp), None, self.stream()),
                    "optmc_european_allreduce")
        return out

    def european_async(self, model, sim, strikes, is_call, moments, terminal=None, aux=None):
        """optmc_european[_aux]_async into the torch tensor ``moments`` (n_contracts*6 f64).
        ``aux`` (3 * n_local f64, one-factor log models): receives W, J and the partner's J."""
        self.n_simulations += 1
        k, c, kp, cp = self._contracts(strikes, is_call)
        tptr = c_void_p(terminal.data_ptr()) if terminal is not None else None
        if aux is not None:
            self._check(self.lib.optmc_european_aux_async(
                self.ctx, ctypes.byref(model), ctypes.byref(sim), k.size, kp, cp,
                c_void_p(moments.data_ptr()), tptr, c_void_p(aux.data_ptr()), self.stream()),
                "optmc_european_aux_async")
            return
        self._check(self.lib.optmc_european_async(self.ctx, ctypes.byref(model),
                                                  ctypes.byref(sim), k.size, kp, cp,
                                                  c_void_p(moments.data_ptr()), tptr,
                                                  self.stream()), "optmc_european_async")

    def european_scen_async(self, model, sim, spot_factors, strike, is_call, moments):
        """optmc_european_scen_async: 2 or 3 bumped-spot scenarios of a SABR kind on one set of
        draws; 6 doubles per scenario into the torch tensor ``moments``."""
        self.n_simulations += 1
        f = np.ascontiguousarray(spot_factors, dtype=np.float64)
        self._check(self.lib.optmc_european_scen_async(
            self.ctx, ctypes.byref(model), ctypes.byref(sim), f.size, f.ctypes.data_as(_dp),
            float(strike), 1 if is_call else 0, c_void_p(moments.data_ptr()), self.stream()),
            "optmc_european_scen_async")

    def sigma_sweep(self, aux, n_local, antithetic, a0, sigma, strikes, is_call):
        """optmc_sigma_sweep_async -> device tensor of n_scenarios * 6 moments."""
        k, c, kp, cp = self._contracts(strikes, is_call)
        a = np.ascontiguousarray(a0, dtype=np.float64)
        sg = np.ascontiguousarray(sigma, dtype=np.float64)
        mom = self.buffer("sigma_moments", k.size * NMOM)
        self._check(self.lib.optmc_sigma_sweep_async(
            self.ctx, c_void_p(aux.data_ptr()), int(n_local), 1 if antithetic else 0, k.size,
            a.ctypes.data_as(_dp), sg.ctypes.data_as(_dp), kp, cp, c_void_p(mom.data_ptr()),
            self.stream()), "optmc_sigma_sweep_async")
        return mom

    def payoff_sweep(self, terminal, n_local, antithetic, strikes, is_call, to_host=True, out=None):
        """optmc_payoff_sweep_async; ``out`` (device tensor, n_contracts * 6 f64) receives the
        moments without any synchronisation, else a cached buffer is used."""
        k, c, kp, cp = self._contracts(strikes, is_call)
        mom = out if out is not None else self.buffer("sweep_moments", k.size * NMOM)
        self._check(self.lib.optmc_payoff_sweep_async(self.ctx, c_void_p(terminal.data_ptr()),
                                                      int(n_local), 1 if antithetic else 0,
                                                      k.size, kp, cp, c_void_p(mom.data_ptr()),
                                                      self.stream()), "optmc_payoff_sweep_async")
        if not to_host or out is not None:
            return mom
        return mom[: k.size * NMOM].cpu().numpy().reshape(k.size, NMOM)

    def local_vol_table(self, vol_surface, n_steps, dt, log_s0, r, q, n_s=1024,
                        half_width=None) -> LocalVolTable:
        """Sample a Dupire surface on the host and place the table in HBM."""
        table, lo, h = sample_local_vol(vol_surface, n_steps, dt, log_s0, r, q, n_s, half_width)
        dev = self.torch.from_numpy(np.ascontiguousarray(table)).to(self.device)
        return LocalVolTable(dev, lo, h)

    # -- single-step terminal samplers --------------------------------------
    def levy_terminal_async(self, levy: OptmcLevy, s0, drift, seed, offset, n_local, strikes,
                            is_call, moments, terminal=None):
        """optmc_levy_terminal_async into the torch tensor ``moments`` (n_contracts*6 f64)."""
        k, c, kp, cp = self._contracts(strikes, is_call)
        tptr = c_void_p(terminal.data_ptr()) if terminal is not None else None
        self._check(self.lib.optmc_levy_terminal_async(
            self.ctx, ctypes.byref(levy), float(s0), float(drift),
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset), int(n_local), k.size, kp, cp,
            c_void_p(moments.data_ptr()), tptr, self.stream()), "optmc_levy_terminal_async")

    # -- fed-draw kernels ---------------------------------------------------
    def fed(self, kind, model: OptmcModel, x0, dt, dw1, dw2=None, counts=None, jumps=None,
            sign=1.0, path_sum_mode=0, want_terminal=True, want_paths=False):
        """
        Run the fed kernel on host (numpy) or device (torch) inputs laid out as
        the reference passes them; returns numpy arrays (terminal, paths).
        """
        torch = self.torch

        def dev(a, dtype):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                return a.to(device=self.device, dtype=dtype).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dtype)

        d1 = dev(dw1, torch.float64)
        n_paths, n_steps = d1.shape
        d2 = dev(dw2, torch.float64)
        cn = dev(counts, torch.int64)
        jb = dev(jumps, torch.float64)
        term = torch.empty(n_paths, dtype=torch.float64, device=self.device) if want_terminal else None
        paths = (torch.empty((n_paths, n_steps + 1), dtype=torch.float64, device=self.device)
                 if want_paths else None)
        ws_bytes = int(self.lib.optmc_fed_workspace_bytes(KIND_CODES[kind], n_paths, n_steps))
        ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device=self.device)
        ptr = lambda t: c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        self._check(self.lib.optmc_fed_async(
            self.ctx, ctypes.byref(model), float(x0), float(dt), n_paths, n_steps, ptr(d1),
            ptr(d2), ptr(cn), ptr(jb), 0 if jb is None else jb.numel(), float(sign),
            int(path_sum_mode), ptr(term), ptr(paths), ptr(ws), ws.numel() * 8, self.stream()),
            "optmc_fed_async")
        return (term.cpu().numpy() if term is not None else None,
                paths.cpu().numpy() if paths is not None else None)

    # -- Longstaff-Schwartz ---------------------------------------------------
    def lsmc_store(self, n_paths, n_steps):
        ld = (int(n_paths) + 1) // 2 * 2
        return self.buffer("lsmc_store", ld * int(n_steps)), ld

    def lsmc_fill_paths(self, model, sim, store, ld):
        self._check(self.lib.optmc_lsmc_paths_async(self.ctx, ctypes.byref(model),
                                                    ctypes.byref(sim), c_void_p(store.data_ptr()),
                                                    ld, self.stream()), "optmc_lsmc_paths_async")

    def lsmc_load(self, paths_dev, n_paths, n_steps, store, ld):
        self._check(self.lib.optmc_lsmc_load_async(self.ctx, c_void_p(paths_dev.data_ptr()),
                                                   n_paths, n_steps, c_void_p(store.data_ptr()),
                                                   ld, self.stream()), "optmc_lsmc_load_async")

    def lsmc(self, store, ld, n_paths, n_steps, strike, is_call, r, dt, degree, peers=False):
        """
        The whole induction in one call: returns (n, sum, sum_sq) of the discounted
        cashflows.  ``peers=True`` (after ``lsmc_peer_attach``): a joint call of all
        attached ranks over their shards; the sums are global.
        """
        if peers or getattr(self, "_peer_key", None) is not None:
            self.set_int("lsmc_use_peers", 1 if peers else 0)
        cash = self.buffer("lsmc_cash", n_paths)
        gram = self.buffer("lsmc_gram", LSMC_GRAM * n_steps)
        out = np.empty(3)
        self._check(self.lib.optmc_lsmc(self.ctx, c_void_p(store.data_ptr()), ld, n_paths,
                                        n_steps, float(strike), 1 if is_call else 0, float(r),
                                        float(dt), int(degree), c_void_p(cash.data_ptr()),
                                        c_void_p(gram.data_ptr()), out.ctypes.data_as(_dp),
                                        self.stream()), "optmc_lsmc")
        return out

    def lsmc_peer_attach(self, group=None):
        """
        Map every rank's exchange buffer into this process (CUDA IPC over NVLink) so
        that ``lsmc`` runs the persistent kernel on all ranks at once, the per-date
        Gram all-reduce fused into it.  Collective over ``group`` (default: world);
        idempotent per group.  Returns (rank, world).
        """
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        key = (id(group), rank, world)
        if getattr(self, "_peer_key", None) == key:
            return rank, world
        if world > MAX_PEERS:
            raise EngineError(f"NVLink exchange supports at most {MAX_PEERS} ranks (one node)")
        mine = (ctypes.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self.lib.optmc_lsmc_peer_export(self.ctx, mine), "optmc_lsmc_peer_export")
        t = self.torch.tensor(list(mine), dtype=self.torch.uint8, device=self.device)
        gathered = [self.torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        blob = bytes(self.torch.cat(gathered).cpu().numpy().tobytes())
        rc = self.lib.optmc_lsmc_peer_attach(self.ctx, rank, world, blob)
        err = self.lib.optmc_last_error(self.ctx).decode() if rc else ""
        # agree on the outcome (this all-reduce is also the barrier: every rank has mapped
        # every buffer before anyone uses one): a rank without peer access must not leave
        # the others waiting in the kernel
        ok = self.torch.tensor([0 if rc else 1], dtype=self.torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.lib.optmc_lsmc_peer_attach(self.ctx, 0, 1, None)
            self._peer_key = None
            raise EngineError("optmc_lsmc_peer_attach: peer mapping failed on at least one rank"
                              + (f" (here: {err})" if err else ""))
        self._peer_key = key
        return rank, world

    def read(self, t, n):
        """The first n doubles of the device tensor ``t`` as a numpy array: through the
        context's pinned buffer (optmc_read_doubles) while they fit, else a torch copy."""
        n = int(n)
        if n > 384:
            return t[:n].cpu().numpy()
        out = np.empty(n)
        self._check(self.lib.optmc_read_doubles(self.ctx, c_void_p(t.data_ptr()), n,
                                                out.ctypes.data_as(_dp), self.stream()),
                    "optmc_read_doubles")
        return out

    PEER_AR_MAX = 1536

    def peer_all_reduce(self, t):
        """Sum the float64 device tensor ``t`` (<= 1536 elements per call, else in chunks) over
        the attached ranks in place: optmc_peer_allreduce_async, a kernel of this library doing
        the exchange over NVLink peer memory on the current stream."""
        n = t.numel()
        for lo in range(0, n, self.PEER_AR_MAX):
            m = min(self.PEER_AR_MAX, n - lo)
            self._check(self.lib.optmc_peer_allreduce_async(
                self.ctx, c_void_p(t.data_ptr() + 8 * lo), m, self.stream()),
                "optmc_peer_allreduce_async")

    def group_all_reduce(self):
        """The all-reduce a technique ends a partitioned pricing with: over NVLink peer memory
        when the process group is NCCL and the ranks' buffers map into each other (attached on
        first use, agreed on by all ranks), else torch.distributed's."""
        import torch.distributed as dist
        mode = getattr(self, "_ar_mode", None)
        if mode is None:
            mode = "dist"
            if dist.get_backend() == "nccl" and os.environ.get("OPTMC_PEER_ALLREDUCE", "1") != "0":
                try:
                    self.lsmc_peer_attach()
                    mode = "peer"
                except EngineError:
                    mode = "dist"
            self._ar_mode = mode
        if mode == "peer":
            self.lsmc_peer_attach()          # idempotent per group: re-attaches after a new group
        return self.peer_all_reduce if mode == "peer" else dist.all_reduce

    def lsmc_peer_detach(self):
        if getattr(self, "_peer_key", None) is not None:
            self._check(self.lib.optmc_lsmc_peer_attach(self.ctx, 0, 1, None),
                        "optmc_lsmc_peer_attach")
            self._peer_key = None

    def lsmc_distributed(self, store, ld, n_paths, n_steps, strike, is_call, r, dt, degree,
                         all_reduce):
        """
        The same induction over this rank's shard of paths with one
        ``all_reduce(tensor)`` of the 24-double power-sum vector per exercise
        date.  ``all_reduce`` sums a device tensor in place
        across ranks; every rank then solves the same normal equations.
        """
        import math

        cash = self.buffer("lsmc_cash", n_paths)
        gram = self.buffer("lsmc_gram", LSMC_GRAM * n_steps)
        disc = math.exp(-r * dt)
        sp, cp, gp = (c_void_p(store.data_ptr()), c_void_p(cash.data_ptr()),
                      c_void_p(gram.data_ptr()))
        for date in range(n_steps - 1, -1, -1):
            self._check(self.lib.optmc_lsmc_step_async(
                self.ctx, sp, ld, n_paths, n_steps, date, float(strike), 1 if is_call else 0,
                disc, int(degree), cp, gp, self.stream()), "optmc_lsmc_step_async")
            all_reduce(gram[date * LSMC_GRAM:(date + 1) * LSMC_GRAM])
        return gram[:3].cpu().numpy()

    # -- diagnostics ----------------------------------------------------------
    def philox_raw(self, seed, n, c2=0, c3=0):
        out = self.torch.empty(4 * n, dtype=self.torch.int32, device=self.device)
        self._check(self.lib.optmc_philox_raw(self.ctx, int(seed), n, c2, c3,
                                              c_void_p(out.data_ptr()), self.stream()),
                    "optmc_philox_raw")
        return out.cpu().numpy().view(np.uint32).reshape(n, 4)

    def normals(self, seed, n, step=0):
        out = self.torch.empty(2 * n, dtype=self.torch.float64, device=self.device)
        self._check(self.lib.optmc_normals(self.ctx, int(seed), n, step, c_void_p(out.data_ptr()),
                                           self.stream()), "optmc_normals")
        return out.cpu().numpy()

    def fp64_peak(self, iters=1 << 16):
        rate, ms = c_double(), c_double()
        self._check(self.lib.optmc_fp64_peak(self.ctx, int(iters), ctypes.byref(rate),
                                             ctypes.byref(ms)), "optmc_fp64_peak")
        return rate.value, ms.value


_engines = {}
_lock = threading.Lock()

# One pricing at a time per process.  The C ABI is not re-entrant per context (optmc.h) and
# Engine.buffer() hands every caller the same cached device buffers, so two techniques priced
# from two Python threads would overwrite each other's terminals / moments.  Every public
# entry point of the techniques and of the fed-kernel wrappers runs under this re-entrant
# lock (greeks call price(); price() may call price_many()).  One process drives one GPU, so
# a process-wide lock serialises exactly the calls that share a context.
PRICING_LOCK = threading.RLock()


def serialized(fn):
    """Decorator: run ``fn`` under PRICING_LOCK."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with PRICING_LOCK:
            return fn(*args, **kwargs)
    return wrapper


def get_engine(device: int | None = None) -> Engine:
    """The process-wide engine for ``device`` (default: torch's current device)."""
    import torch

    if not torch.cuda.is_available():
        raise EngineError("no CUDA device visible: the optpricing_b200 engine needs a B200 "
                          "(there is no CPU fallback)")
    idx = torch.cuda.current_device() if device is None else int(device)
    with _lock:
        eng = _engines.get(idx)
        if eng is None:
            eng = _engines[idx] = Engine(idx)
    return eng
