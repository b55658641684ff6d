#!/usr/bin/env python
"""
bench.py -- path-steps/sec of the Monte Carlo hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c1|c4] [--scaling weak|strong]

One "step" = one pricing of the workload's contract: the whole hot path
(Philox draws -> SDE steps -> payoff moments [-> NCCL all-reduce]) over one
batch of synthetic paths.  Default workload = BASELINE configs[1] (C2): Heston
European call, 2^24 paths x 252 steps, antithetic, fp64, reference-compatible
correlation.  With N > 1 (torchrun, one rank per GPU) the draws are partitioned
by rank; "weak" (default) keeps 2^24 paths per GPU, "strong" keeps 2^24 total.

Printed JSON (rank 0, one line):
  value     device-timed whole-job path-steps/s (CUDA events, max over ranks);
            the kernel has no HBM inputs -- its inputs are ~400 B of parameters
  e2e       the same metric through technique.price() (the reference-facing
            Python API -> ctypes C ABI), wall clock, parameters from host
            memory in, 48 B of moments read back every step
  roofline  fp64-pipe roofline of the dominant kernel: achieved
            = path-steps/s x fp64-pipe instructions per path-step (static SASS
            count, profiles/), peak = DFMA/s measured live by optmc_fp64_peak
  cpu_baseline  the oracle port of the reference's CPU path timed on the host

--impl reference times the oracle port (the reference is Python+numba and does
not travel to the GPU box) with all host threads, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# fp64-pipe thread-instructions per path-step of the dominant kernel, counted
# from the SASS of its step loop (tools/sass_stats.py; see profiles/ and
# DESIGN.md "Roofline").  Pair-step count / 2 for the antithetic kernels.
I64_PER_PATH_STEP = {"heston": 16.5, "bsm": 5.5}
# Algorithmic flops per path-step (antithetic pairs share the draws)
FLOPS_PER_PATH_STEP = {"heston": 28.0, "bsm": 6.0}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
# `ncu --set full` captures summarised in profiles/r01_ncu_heston_summary.md (r1f) and
# profiles/r01_ncu_lsmc_summary.md (persistent): the European kernel reads its tables and code only;
# the LSMC induction streams the 1.68 GB path store once (the 6.7 GB of algorithmic cashflow / row
# re-reads are L2 hits at 2^22 paths).
NCU_TRAFFIC_BYTES = {"heston_c2": 52_736 + 0, "lsmc_c4": 1_729_129_000 + 123_759_616}

HESTON = {"v0": 0.04, "kappa": 2.0, "theta": 0.04, "rho": -0.7, "vol_of_vol": 0.5}
WORKLOADS = {
    "c2": dict(kind="heston", params=HESTON, n_paths=1 << 24, n_steps=252, S0=100.0, K=100.0,
               T=1.0, r=0.05, q=0.0, name="C2 Heston European call 2^24 paths x 252 steps, "
               "antithetic, fp64, reference correlation"),
    "c1": dict(kind="bsm", params={"sigma": 0.2}, n_paths=100_000, n_steps=252, S0=100.0,
               K=100.0, T=1.0, r=0.05, q=0.0,
               name="C1 BSM European call 100k paths x 252 steps, antithetic"),
}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.all_rows, self.rows, self.proc = [], [], None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.all_rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        """Start of the timed region: only samples taken from here on count."""
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.t1 = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        # a sample printed at time t describes the interval just before t
        self.rows = [r for t, r in self.all_rows if t0 + 0.01 <= t <= self.t1 + 0.03]
        if not self.rows:      # timed region shorter than the sampling period: nearest samples
            self.rows = [r for t, r in self.all_rows if t >= t0 - 0.05][:3] or [
                r for _, r in self.all_rows[-2:]]
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows),
                "reasons": reasons}


# ---------------------------------------------------------------------------
# CPU side: the oracle port of the reference's path, all host threads
# ---------------------------------------------------------------------------
def cpu_port_rate(w, threads, chunks_per_thread=2, chunk_paths=1 << 15, repeats=1):
    """path-steps/s of the oracle restatement of MonteCarloTechnique.price
    (monte_carlo.py:65-257: numpy draws + C step loop + payoff mean) with
    `threads` workers each pricing independent chunks of the same contract."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from oracle import reference_mc as orc
    orc.lib()

    def work(i):
        rng = np.random.default_rng([42, i])
        return orc.price_european(w["kind"], w["params"], w["S0"], w["K"], w["T"], w["r"],
                                  w["q"], True, chunk_paths, w["n_steps"], True, rng,
                                  jump_rs=np.random.RandomState(i))

    n_chunks = threads * chunks_per_thread
    best, prices = float("inf"), None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            prices = list(ex.map(work, range(n_chunks)))
            best = min(best, time.perf_counter() - t0)
    units = n_chunks * chunk_paths * w["n_steps"]
    return units / best, best, float(sum(prices) / len(prices)), n_chunks * chunk_paths


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 64)
    for _ in range(args.warmup):
        cpu_port_rate(w, threads, chunks_per_thread=1, chunk_paths=1 << 12)
    times, sample_paths, price = [], 0, None
    for _ in range(args.steps):
        rate, secs, price, sample_paths = cpu_port_rate(w, threads)
        times.append(secs)
    total = sum(times)
    value = args.steps * sample_paths * w["n_steps"] / total
    sample = (f"{sample_paths} paths x {w['n_steps']} steps per step ({threads} threads x 2 "
              f"chunks of 32768 paths), oracle port of monte_carlo.py:65-257 "
              f"(numpy PCG64 draws + C step loop); price {price:.4f}")
    line = {
        "impl": "reference", "metric": "path-steps/sec (fp64)", "value": value,
        "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
def run_ours(args, w):
    import ctypes

    import numpy as np
    import torch

    import optpricing_b200 as ob
    from optpricing_b200 import _engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = _engine.get_engine(local)
    dev = eng.device

    n_job = w["n_paths"] * (world if args.scaling == "weak" else 1)
    n_draws = n_job // 2
    off, cnt = ob.techniques.monte_carlo.shard_draws(n_draws, rank, world)
    mdl = _engine.make_model(w["kind"], w["params"], w["r"], w["q"])
    moments = torch.zeros(_engine.NMOM, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    units_per_step = n_job * w["n_steps"]

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def device_step(seed):
        sim = eng.make_sim(w["S0"], w["T"], n_draws, w["n_steps"], True, seed,
                           _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
        eng.european_async(mdl, sim, [w["K"]], [1], moments)
        if dist:
            dist.all_reduce(moments)

    # ---- device-timed value -------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None    # started early: nvidia-smi is slow to start
    for i in range(args.warmup):
        device_step(1000 + i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    launches0 = eng.launch_count()
    barrier()
    if sampler:
        sampler.begin()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                   # L2 flush, outside the step's events
        ev[i][0].record()
        kev[i][0].record()
        sim = eng.make_sim(w["S0"], w["T"], n_draws, w["n_steps"], True, 2000 + i,
                           _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
        eng.european_async(mdl, sim, [w["K"]], [1], moments)
        kev[i][1].record()
        if dist:
            dist.all_reduce(moments)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kern_total = torch.tensor([sum(kern_ms)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_total, op=dist.ReduceOp.MAX)
    total_ms, kern_total = float(total_ms.item()), float(kern_total.item())
    value = args.steps * units_per_step / (total_ms * 1e-3)
    stats_dev = ob.techniques.monte_carlo.summarize(
        moments.cpu().numpy(), w["r"], w["T"],
        ob.techniques.monte_carlo.control_mean(w["kind"], w["params"], w["S0"], w["r"], w["q"],
                                               w["T"], w["n_steps"]))

    # ---- end to end through the public API ----------------------------------
    Model = {"heston": ob.HestonModel, "bsm": ob.BSMModel}[w["kind"]]
    option = ob.Option(strike=w["K"], maturity=w["T"], option_type=ob.OptionType.CALL)
    stock, rate, model = ob.Stock(spot=w["S0"], dividend=w["q"]), ob.Rate(rate=w["r"]), Model(w["params"])
    tech = ob.MonteCarloTechnique(n_paths=n_job, n_steps=w["n_steps"], antithetic=True, seed=42)
    for _ in range(args.warmup):
        tech.price(option, stock, model, rate)
    barrier()
    t0 = time.perf_counter()
    price = None
    for _ in range(args.steps):
        price = tech.price(option, stock, model, rate).price   # returns a host float each step
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = args.steps * units_per_step / float(e2e_s.item())
    h2d = ctypes.sizeof(_engine.OptmcModel) + ctypes.sizeof(_engine.OptmcSim) + 8 + 4
    d2h = 8 * _engine.NMOM

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0's GPU) ----------------------
    peak_burst, _ = eng.fp64_peak(1 << 17)      # ~20 ms, like one step
    peak_sust, _ = eng.fp64_peak(1 << 21)       # ~0.3 s
    peak = max(peak_burst, peak_sust)
    local_units = cnt * 2 * w["n_steps"]
    kern_s = kern_total * 1e-3 / args.steps
    i64 = I64_PER_PATH_STEP[w["kind"]]
    achieved = local_units * i64 / kern_s
    roofline = {
        "bound": "fp64", "kernel": f"k_european<{w['kind']}, antithetic, hand-written normals>",
        "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G fp64-pipe instr/s",
        "frac": achieved / peak,
        "traffic": NCU_TRAFFIC_BYTES["heston_c2"] if args.workload == "c2" and world == 1 else None,
        "traffic_note": "bytes of DRAM read+written per launch, ncu capture of this workload on one "
                        "GPU (profiles/r01_ncu_heston_summary.md); the kernel has no HBM operands",
        "per_unit": f"{i64} fp64-pipe instr per path-step (static SASS count of the step loop)",
        "peak_source": "optmc_fp64_peak DFMA loop measured in this run (MEASURED_PEAKS.json has "
                       "no fp64 figure); burst %.1f / sustained %.1f G/s" % (peak_burst / 1e9,
                                                                             peak_sust / 1e9),
        "alg_tflops": local_units * FLOPS_PER_PATH_STEP[w["kind"]] / kern_s / 1e12,
        "alg_flops_per_path_step": FLOPS_PER_PATH_STEP[w["kind"]],
        "kernel_ms": kern_s * 1e3,
        "hbm_note": "no HBM traffic in the step loop: state in registers, 40 B per block out",
    }

    # ---- CPU baseline (oracle port), bounded sample ---------------------------
    threads = min(os.cpu_count() or 1, 64)
    cpu_port_rate(w, threads, chunks_per_thread=1, chunk_paths=1 << 12)       # warm
    rate_all, secs, cpu_price, sample_paths = cpu_port_rate(w, threads, repeats=2)
    rate_1, secs1, _, sp1 = cpu_port_rate(w, 1, chunks_per_thread=4, repeats=1)
    cpu = {"value": rate_all, "unit": "path-steps/s", "cores": threads, "kind": "port",
           "sample": f"{sample_paths} paths x {w['n_steps']} steps in {secs:.2f} s "
                     f"(oracle port of monte_carlo.py:65-257, numpy draws + C step loop); "
                     f"single thread: {rate_1:.3e} path-steps/s on {sp1} paths",
           "single_thread_value": rate_1, "price": cpu_price}

    line = {
        "metric": "path-steps/sec (fp64)", "value": value, "unit": "path-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "paths_per_step": n_job, "n_steps": w["n_steps"],
                   "partition": f"draws sharded over {world} rank(s), one all-reduce of 6 doubles",
                   "l2": "256 MiB written between timed steps (outside the per-step events); "
                         "the kernel reads no HBM input"},
        "e2e": {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "MonteCarloTechnique.price", "price": price,
                "stderr": tech.last_stats.get("stderr")},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "check": {"price": stats_dev["price"], "stderr": stats_dev["stderr"],
                  "control_mean": stats_dev["control_mean"],
                  "control_expected": stats_dev["control_expected"],
                  "wall_s_timed_region": t_wall},
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# secondary workloads (BASELINE configs 3-5): through the public API only
# ---------------------------------------------------------------------------
def run_secondary(args, name):
    import numpy as np
    import torch

    import optpricing_b200 as ob
    from optpricing_b200 import _engine

    # Under torchrun the techniques shard their paths over the ranks themselves (strong
    # scaling: the workload's path count is the whole job's); rank 0 prints the line.
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = _engine.get_engine(local)
    stock, rate = ob.Stock(spot=100.0, dividend=0.01), ob.Rate(rate=0.05)
    extra = {}

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
    if name == "c3":
        model = ob.BatesModel()
        strikes = np.linspace(70, 130, 64)
        mats = (0.25, 0.5, 1.0, 2.0)
        tech = ob.MonteCarloTechnique(n_paths=1 << 22, steps_per_year=252, seed=7)
        opts = {T: [ob.Option(float(k), T, ob.OptionType.CALL) for k in strikes] for T in mats}
        grid = [o for T in mats for o in opts[T]]
        units = sum((1 << 22) * int(252 * T) for T in mats)
        label = ("C3 Bates European strike sweep, 64 strikes x 4 maturities, 2^22 paths per maturity, "
                 "n_steps = 252 T, one simulation per maturity, ONE price_many call for the grid")

        def step():
            return tech.price_many(grid, stock, model, rate)
        extra["contracts_per_step"] = 256
    elif name == "c4":
        model = ob.BSMModel({"sigma": 0.2})
        aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
        tech = ob.AmericanMonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=42, lsm_degree=3)
        units = (1 << 22) * 50
        label = "C4 American put Longstaff-Schwartz (GBM), 2^22 paths x 50 exercise dates, cubic basis"

        def step():
            return tech.price(aput, stock, model, rate).price
    elif name == "c5":
        call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
        models = (ob.SABRJumpModel(), ob.KouModel())
        techs = [ob.MonteCarloTechnique(n_paths=1 << 24, n_steps=365, seed=11, control_variate=True)
                 for _ in models]
        # per step: price + delta + vega of both models on common random numbers.
        # SABR-jump: 1 + 2 simulations (bump and re-price; its vega is NaN as in the reference,
        # greek_mixin.py:156-157).  Kou: 1 + 1 + 1 -- its bumped prices are read off one
        # simulation each (spot scaling for delta, stored (W, J) for vega), which is what the
        # reference's 2 + 2 re-pricings on common random numbers compute.  The value counts
        # the simulations actually run (engine counter), each 2^24 paths x 365 steps.
        units = None
        label = ("C5 SABR-with-jumps and Kou European call, 2^24 paths x 365 steps, control variate, "
                 "price + delta + vega on common random numbers")

        def step():
            out = []
            for t, m in zip(techs, models):
                out += [t.price(call, stock, m, rate).price, t.delta(call, stock, m, rate),
                        t.vega(call, stock, m, rate)]
            return out
    else:
        raise SystemExit(name)
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    launches0, sims0 = eng.launch_count(), eng.n_simulations
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(args.steps):
        out = step()
    b.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = a.elapsed_time(b)
    if units is None:          # c5: simulations actually run x paths x steps
        sims = (eng.n_simulations - sims0) // args.steps
        units = sims * (1 << 24) * 365
        extra["simulations_per_step"] = sims
    if dist:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0].item()), float(t[1].item())
        if rank != 0:
            dist.destroy_process_group()
            return
    line = {"metric": "path-steps/sec (fp64)", "value": args.steps * units / (dev_ms * 1e-3),
            "unit": "path-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "l2": "path store / terminal buffers exceed L2 or no HBM "
                                                "input; no flush between steps", **extra},
            "e2e": {"value": args.steps * units / wall, "unit": "path-steps/s",
                    "h2d_bytes_per_step": 212, "d2h_bytes_per_step": 48, "api": "technique.price"},
            "gpu_launches": eng.launch_count() - launches0,
            "result_sample": (float(np.ravel(out)[0]) if not isinstance(out, float) else out)}
    if name == "c3" and world == 1:
        # SURVEY 8 d2 asks for both forms: the batched sweep (the line's value) and the drop-in
        # loop of 256 independent price() calls, each simulating its own 2^22 paths
        t = {T: ob.MonteCarloTechnique(n_paths=1 << 22, n_steps=int(252 * T), seed=7) for T in mats}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop_prices = [t[T].price(o, stock, model, rate).price for T in mats for o in opts[T]]
        torch.cuda.synchronize()
        loop_s = time.perf_counter() - t0
        line["config"]["drop_in_loop"] = {
            "calls": len(loop_prices), "seconds": loop_s,
            "path_steps_per_s": 64 * units / loop_s,
            "note": "256 technique.price() calls, one simulation of 2^22 paths each"}
    if name == "c5":
        line["config"]["results"] = dict(zip(
            ("sabr_jump_price", "sabr_jump_delta", "sabr_jump_vega", "kou_price", "kou_delta",
             "kou_vega"), (None if (isinstance(x, float) and math.isnan(x)) else float(x) for x in out)))
    if name == "c4" and world == 1:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        store, ld = eng.lsmc_store(1 << 22, 50)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.lsmc(store, ld, 1 << 22, 50, 110.0, False, 0.05, 1.0 / 50, 3)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        ach = (1 << 22) * 50 * 32 / (best * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "k_lsmc_persistent<3>: 50 exercise dates in one "
                                                      "launch (induction only)",
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": NCU_TRAFFIC_BYTES["lsmc_c4"],
                            "traffic_note": "DRAM bytes per launch (ncu, profiles/r01_ncu_lsmc_summary.md):"
                                            " the 1.68 GB store once + cashflow write-backs; the other "
                                            "3/4 of the 6.7 GB algorithmic bytes are L2 hits, so the "
                                            "launch is bound by fp64 issue, not by HBM",
                            "per_unit": "32 B per path-date (SURVEY 8 d4)",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                            "kernel_ms": best}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c4", "c5"])
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload in ("c3", "c4", "c5"):
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined for the c1/c2 workloads")
        return run_secondary(args, args.workload)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
