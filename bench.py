#!/usr/bin/env python
"""
bench.py -- path-steps/sec of the Monte Carlo hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c1|c3|c4|c5] [--scaling strong|weak] [--no-secondary]

One "step" = one pricing of the workload's contract: the whole hot path
(Philox draws -> SDE steps -> payoff moments [-> NCCL all-reduce]) over one
batch of synthetic paths.  Default workload = BASELINE configs[1] (C2): Heston
European call, 2^24 paths x 252 steps, antithetic, fp64, reference-compatible
correlation.  With N > 1 (torchrun, one rank per GPU) the draws are partitioned by
rank.  Default scaling is STRONG: the job stays 2^24 paths whatever N is -- the
north star's ">= 7x from 1 to 8 GPUs" is value(N = 8) / value(N = 1) of this
line; the weak-scaling figure (2^24 paths per GPU) rides along in `weak`.

Printed JSON (rank 0, one line):
  value     device-timed whole-job path-steps/s (CUDA events, max over ranks);
            the kernel has no HBM inputs -- its inputs are ~400 B of parameters
  e2e       the same metric through technique.price() (the reference-facing
            Python API -> ctypes C ABI), wall clock, parameters from host
            memory in, 48 B of moments read back every step
  roofline  fp64-pipe roofline of the dominant kernel: achieved
            = path-steps/s x fp64-pipe instructions per path-step (static SASS
            count, profiles/), peak = DFMA/s measured live by optmc_fp64_peak;
            `issue`: instructions per path-step / cycles per path-step per SM
            sub-partition, and the register-operand model that bounds it
  cpu_baseline  the oracle port of the reference's CPU path timed on the host
            (all threads, and one thread), and `reference_numba`: the reference
            itself (baseline/_ref, numba, single thread) timed in the same run
  secondary BASELINE configs C1, C3, C4, C5 on the same GPUs: ms per step, value,
            roofline (C4: HBM), cpu_baseline -- what round 1 only had in builder logs
  parity_multi (N > 1) distributed result == single-GPU result, C2 moments and
            C4 price through both exchanges (NVLink in-kernel, NCCL per date)

--impl reference times the oracle port (all host threads) on a bounded sample and
adds the numba reference's single-thread figure; the reference is Python + numba,
single-threaded by construction, and cannot allocate C2 at full size.
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

# Static SASS counts of the dominant kernel's step loop per PATH-step (one antithetic
# pair-step / 2), tools/sass_stats.py on k_european_wide<kind, antithetic> (profiles/r02_*):
#   i64    fp64-pipe instructions (DFMA DMUL DADD)
#   instr  all instructions
#   reads  32-bit vector-register operand reads (B200: 2 per cycle per SM sub-partition,
#          profiles/r01_pipe_probe3.log -- a DFMA with three distinct register operands
#          takes 3 cycles, not 2)
SASS = {"heston": {"i64": 13.0, "instr": 38.75, "reads": 88.0},
        "bsm": {"i64": 3.75, "instr": 15.5, "reads": 31.6}}
# Algorithmic flops per path-step (antithetic pairs share the draws)
FLOPS_PER_PATH_STEP = {"heston": 28.0, "bsm": 6.0}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
# `ncu --set full` captures summarised in profiles/ (r02_ncu_heston_summary.md,
# r02_ncu_lsmc_summary.md).  The European kernel reads its tables and code only.
NCU_TRAFFIC_BYTES = {"heston_c2": 48_896 + 0, "lsmc_c4_induction": 1_778_023_000 + 178_389_000,
                     "lsmc_c4_fill": 1_677_721_600}

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
    """nvidia-smi clocks/throttle reasons sampled every 20 ms during the timed region."""
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
# CPU side: the oracle port of the reference's path, and the reference itself
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


def cpu_port_lsmc_rate(threads, n_paths=1 << 17, n_steps=50, degree=3):
    """path-steps/s of the oracle restatement of AmericanMonteCarloTechnique.price (C4's
    contract: numpy draws + C path kernel + numpy Longstaff-Schwartz), `threads` independent
    pricings side by side (the reference has no way to split one pricing)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from oracle import reference_mc as orc
    orc.lib()

    def work(i):
        return orc.price_american("bsm", {"sigma": 0.2}, 100.0, 110.0, 1.0, 0.05, 0.01, False,
                                  n_paths, n_steps, True, np.random.default_rng([42, i]),
                                  degree=degree)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        t0 = time.perf_counter()
        prices = list(ex.map(work, range(threads)))
        secs = time.perf_counter() - t0
    return threads * n_paths * n_steps / secs, secs, float(sum(prices) / len(prices))


def reference_numba(name, timeout=240):
    """The unmodified reference (baseline/_ref, numba, single thread) on this host."""
    script = os.path.join(ROOT, "tools", "time_reference_numba.py")
    try:
        r = subprocess.run([sys.executable, script, name], capture_output=True, text=True,
                           timeout=timeout, env=dict(os.environ, NUMBA_CACHE_DIR="/tmp/numba_cache"))
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": (r.stderr or r.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout} s"}
    except Exception as e:  # noqa: BLE001 - a baseline that cannot run must not cost the bench line
        return {"unavailable": repr(e)[:300]}


def cpu_baseline_c2(w, with_numba=True):
    threads = min(os.cpu_count() or 1, 64)
    cpu_port_rate(w, threads, chunks_per_thread=1, chunk_paths=1 << 12)       # warm
    rate_all, secs, cpu_price, sample_paths = cpu_port_rate(w, threads, repeats=2)
    rate_1, _, _, sp1 = cpu_port_rate(w, 1, chunks_per_thread=4, repeats=1)
    cpu = {"value": rate_all, "unit": "path-steps/s", "cores": threads, "kind": "port",
           "sample": f"{sample_paths} paths x {w['n_steps']} steps in {secs:.2f} s "
                     f"(oracle port of monte_carlo.py:65-257, numpy draws + C step loop); "
                     f"single thread: {rate_1:.3e} path-steps/s on {sp1} paths",
           "single_thread_value": rate_1, "price": cpu_price, "nproc": os.cpu_count()}
    if with_numba:
        cpu["reference_numba"] = reference_numba("c2" if w["kind"] == "heston" else "c1")
    return cpu


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
    rate_1, _, _, _ = cpu_port_rate(w, 1, chunks_per_thread=2, repeats=1)
    sample = (f"{sample_paths} paths x {w['n_steps']} steps per step ({threads} threads x 2 "
              f"chunks of 32768 paths), oracle port of monte_carlo.py:65-257 "
              f"(numpy PCG64 draws + C step loop); price {price:.4f}")
    line = {
        "impl": "reference", "metric": "path-steps/sec (fp64)", "value": value,
        "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": threads,
                         "kind": "port", "sample": sample, "single_thread_value": rate_1,
                         "nproc": os.cpu_count(),
                         "reference_numba": reference_numba(
                             "c2" if w["kind"] == "heston" else "c1")},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch

        from optpricing_b200 import _engine
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.eng = _engine.get_engine(self.local)
        self.dev = self.eng.device
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def c2_device(cx, args, w, scaling, steps, sampler=None):
    """Device-timed steps of the C2 hot path: kernel + finishing pass [+ all-reduce]."""
    import optpricing_b200 as ob
    from optpricing_b200 import _engine
    torch, eng = cx.torch, cx.eng
    n_job = w["n_paths"] * (cx.world if scaling == "weak" else 1)
    n_draws = n_job // 2
    off, cnt = ob.techniques.monte_carlo.shard_draws(n_draws, cx.rank, cx.world)
    mdl = _engine.make_model(w["kind"], w["params"], w["r"], w["q"])
    moments = torch.zeros(_engine.NMOM, dtype=torch.float64, device=cx.dev)

    def step(seed, kev=None):
        sim = eng.make_sim(w["S0"], w["T"], n_draws, w["n_steps"], True, seed,
                           _engine.CORR_REFERENCE_EUROPEAN, off, cnt)
        eng.european_async(mdl, sim, [w["K"]], [1], moments)
        if kev is not None:
            kev.record()
        if cx.dist:
            all_reduce(moments)

    # what MonteCarloTechnique.price does after the kernel: the engine's own all-reduce kernel
    # over NVLink peer memory (csrc/lsmc.cuh: k_peer_allreduce) on an NCCL group
    all_reduce = eng.group_all_reduce() if cx.dist else None
    for i in range(max(args.warmup, 3)):
        step(1000 + i)
    cx.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = eng.launch_count()
    cx.barrier()
    if sampler:
        sampler.begin()
    t0 = time.perf_counter()
    for i in range(steps):
        cx.flush.zero_()                                 # L2 flush, outside the step's events
        ev[i][0].record()
        step(2000 + i, ev[i][1])
        ev[i][2].record()
    cx.barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    total_ms, kern_ms = cx.max_over_ranks(total_ms, kern_ms)
    return {"n_job": n_job, "cnt": cnt, "ms_per_step": total_ms / steps,
            "kernel_ms": kern_ms / steps, "value": steps * n_job * w["n_steps"] / (total_ms * 1e-3),
            "launches": launches, "clocks": clocks, "wall": wall,
            "moments": moments.cpu().numpy()}


def c2_e2e(cx, args, w, scaling, steps):
    import optpricing_b200 as ob
    n_job = w["n_paths"] * (cx.world if scaling == "weak" else 1)
    Model = {"heston": ob.HestonModel, "bsm": ob.BSMModel}[w["kind"]]
    option = ob.Option(strike=w["K"], maturity=w["T"], option_type=ob.OptionType.CALL)
    stock, rate, model = ob.Stock(spot=w["S0"], dividend=w["q"]), ob.Rate(rate=w["r"]), Model(w["params"])
    tech = ob.MonteCarloTechnique(n_paths=n_job, n_steps=w["n_steps"], antithetic=True, seed=42)
    for _ in range(max(args.warmup, 3)):
        tech.price(option, stock, model, rate)
    cx.barrier()
    t0 = time.perf_counter()
    price = None
    for _ in range(steps):
        price = tech.price(option, stock, model, rate).price   # returns a host float each step
    cx.barrier()
    (secs,) = cx.max_over_ranks(time.perf_counter() - t0)
    return {"value": steps * n_job * w["n_steps"] / secs, "price": price,
            "stderr": tech.last_stats.get("stderr"), "seed": tech.last_stats.get("seed"),
            "tech": tech, "args": (option, stock, model, rate)}


def fp64_roofline(cx, w, cnt, kernel_ms, clocks, world, is_c2):
    eng = cx.eng
    peak_burst, _ = eng.fp64_peak(1 << 17)      # ~20 ms, like one step
    peak_sust, _ = eng.fp64_peak(1 << 21)       # ~0.3 s
    peak = max(peak_burst, peak_sust)
    local_units = cnt * 2 * w["n_steps"]
    kern_s = kernel_ms * 1e-3
    sass = SASS[w["kind"]]
    achieved = local_units * sass["i64"] / kern_s
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sm_count = cx.torch.cuda.get_device_properties(cx.dev).multi_processor_count
    # cycles one SM sub-partition spends per warp-level path-step (32 paths advance one step)
    cyc = kern_s * mhz * 1e6 / (local_units / 32.0 / (sm_count * 4))
    return {
        "bound": "fp64", "kernel": f"k_european_wide<{w['kind']}, antithetic>: one 768-thread block "
                                   "per SM on lane-replicated (conflict-free) tables",
        "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G fp64-pipe instr/s",
        "frac": achieved / peak,
        "traffic": NCU_TRAFFIC_BYTES["heston_c2"] if is_c2 and world == 1 else None,
        "traffic_note": "bytes of DRAM read+written per launch, ncu capture of this workload on one "
                        "GPU (profiles/r02_ncu_heston_summary.md); the kernel has no HBM operands",
        "per_unit": f"{sass['i64']} fp64-pipe instr per path-step (static SASS count of the step loop)",
        "peak_source": "optmc_fp64_peak DFMA loop measured in this run (MEASURED_PEAKS.json has "
                       "no fp64 figure); burst %.1f / sustained %.1f G/s" % (peak_burst / 1e9,
                                                                             peak_sust / 1e9),
        "alg_tflops": local_units * FLOPS_PER_PATH_STEP[w["kind"]] / kern_s / 1e12,
        "alg_flops_per_path_step": FLOPS_PER_PATH_STEP[w["kind"]],
        "kernel_ms": kernel_ms,
        "issue": {
            "instr_per_path_step": sass["instr"], "cycles_per_path_step_per_smsp": cyc,
            "frac": sass["instr"] / cyc,
            "operand_reads_per_path_step": sass["reads"],
            "operand_bound_cycles": sass["reads"] / 2.0,
            "operand_frac": sass["reads"] / 2.0 / cyc,
            "note": "one issue slot and two 32-bit register-operand reads per cycle per SM "
                    "sub-partition (profiles/r01_pipe_probe3.log): a DFMA with three distinct "
                    "register operands occupies the operand path for 3 cycles, so this "
                    "instruction mix cannot reach 100 % of the 2-cycle fp64-pipe roofline",
            "sm_mhz_used": mhz},
        "hbm_note": "no HBM traffic in the step loop: state in registers, 40 B per block out",
    }


# ---------------------------------------------------------------------------
# secondary workloads (BASELINE configs 1, 3, 4, 5): through the public API
# ---------------------------------------------------------------------------
def secondary(cx, name, steps, with_cpu=True):
    import numpy as np

    import optpricing_b200 as ob
    torch, eng = cx.torch, cx.eng
    stock, rate = ob.Stock(spot=100.0, dividend=0.01), ob.Rate(rate=0.05)
    extra, units = {}, None
    if name == "c1":
        w = WORKLOADS["c1"]
        model = ob.BSMModel(w["params"])
        call = ob.Option(strike=100.0, maturity=1.0, option_type=ob.OptionType.CALL)
        stock = ob.Stock(spot=100.0)
        tech = ob.MonteCarloTechnique(n_paths=w["n_paths"], n_steps=252, seed=42)
        units, label = w["n_paths"] * 252, w["name"]

        def step():
            return tech.price(call, stock, model, rate).price
    elif name == "c3":
        model = ob.BatesModel()
        strikes = np.linspace(70, 130, 64)
        mats = (0.25, 0.5, 1.0, 2.0)
        tech = ob.MonteCarloTechnique(n_paths=1 << 22, steps_per_year=252, seed=7)
        grid = [ob.Option(float(k), T, ob.OptionType.CALL) for T in mats for k in strikes]
        units = sum((1 << 22) * int(252 * T) for T in mats)
        label = ("C3 Bates European strike sweep, 64 strikes x 4 maturities, 2^22 paths per maturity, "
                 "n_steps = 252 T, one simulation per maturity, ONE price_many call for the grid")
        extra["contracts_per_step"] = 256

        def step():
            return tech.price_many(grid, stock, model, rate)
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
        # per step: price + delta + vega of both models on common random numbers, ONE
        # price_with_greeks call per model = ONE simulation per model: SABR-jump carries the
        # spot scenarios (up, base, down) side by side on shared draws (its vega is NaN as in
        # the reference, greek_mixin.py:156-157); Kou's bumped prices are strike rescalings
        # and a sigma sweep of the one simulation.  The value counts simulations actually run
        # (engine counter), each 2^24 paths x 365 steps -- round 1 ran 6 per step.
        label = ("C5 SABR-with-jumps and Kou European call, 2^24 paths x 365 steps, control variate, "
                 "price + delta + vega on common random numbers (price_with_greeks)")

        def step():
            out = []
            for t, m in zip(techs, models):
                res = t.price_with_greeks(call, stock, m, rate, greeks=("delta", "vega"))
                out += [res.price, res.greeks["delta"], res.greeks["vega"]]
            return out
    else:
        raise SystemExit(name)
    for _ in range(3):
        out = step()
    cx.barrier()
    launches0, sims0 = eng.launch_count(), eng.n_simulations
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        out = step()
    b.record()
    cx.barrier()
    wall = time.perf_counter() - t0
    dev_ms, wall = cx.max_over_ranks(a.elapsed_time(b), wall)
    if units is None:          # c5: simulations actually run x paths x steps
        sims = (eng.n_simulations - sims0) // steps
        units = sims * (1 << 24) * 365
        extra["simulations_per_step"] = sims
    res = {"workload": label, "ms_per_step": dev_ms / steps, "value": steps * units / (dev_ms * 1e-3),
           "unit": "path-steps/s", "steps": steps, "scaling": "strong",
           "e2e": {"value": steps * units / wall, "unit": "path-steps/s", "api": "technique.price"},
           "gpu_launches": eng.launch_count() - launches0,
           "result_sample": (float(np.ravel(out)[0]) if not isinstance(out, float) else out),
           **extra}
    if name == "c5":
        res["results"] = dict(zip(
            ("sabr_jump_price", "sabr_jump_delta", "sabr_jump_vega", "kou_price", "kou_delta",
             "kou_vega"), (None if (isinstance(x, float) and math.isnan(x)) else float(x) for x in out)))
    if cx.rank != 0:
        return res
    threads = min(os.cpu_count() or 1, 64)
    if name == "c4":
        res["roofline"] = c4_roofline(cx, dev_ms / steps) if cx.world == 1 else None
        if with_cpu:
            rate_all, secs, price = cpu_port_lsmc_rate(threads)
            rate_1, _, _ = cpu_port_lsmc_rate(1)
            res["cpu_baseline"] = {
                "value": rate_all, "unit": "path-steps/s", "cores": threads, "kind": "port",
                "sample": f"{threads} independent pricings of 2^17 paths x 50 dates, cubic basis, "
                          f"in {secs:.2f} s (oracle port of american_monte_carlo.py:52-207 + "
                          f"american_mc_kernels.py:36-84); price {price:.4f}",
                "single_thread_value": rate_1, "nproc": os.cpu_count(),
                "reference_numba": reference_numba("c4")}
    elif with_cpu and name in ("c3", "c5"):
        kinds = {"c3": [("bates", ob.BatesModel().params, 252)],
                 "c5": [("sabr_jump", ob.SABRJumpModel().params, 365),
                        ("kou", ob.KouModel().params, 365)]}[name]
        tot_units = tot_secs = 0.0
        for kind, p, m in kinds:
            w = dict(kind=kind, params=p, S0=100.0, K=100.0, T=1.0, r=0.05, q=0.01, n_steps=m)
            rate_all, secs, _, sp = cpu_port_rate(w, threads, chunks_per_thread=1)
            tot_units += sp * m
            tot_secs += secs
        res["cpu_baseline"] = {
            "value": tot_units / tot_secs, "unit": "path-steps/s", "cores": threads, "kind": "port",
            "sample": f"one simulation of {threads} x 32768 paths per model ({', '.join(k for k, _, _ in kinds)}), "
                      f"oracle port of monte_carlo.py:65-257 on all host threads; the reference "
                      f"re-simulates per strike / per bumped pricing, the rate per path-step is the same",
            "nproc": os.cpu_count()}
    return res


def c4_roofline(cx, price_ms):
    """C4 is bound by HBM only in principle: what bounds it is stated, not assumed."""
    eng, torch = cx.eng, cx.torch
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
    store_bytes = (1 << 22) * 50 * 8
    traffic = NCU_TRAFFIC_BYTES["lsmc_c4_induction"]
    ach = traffic / (best * 1e-3) / 1e9
    floor_ms = 2 * store_bytes / (peak * 1e9) * 1e3
    return {"bound": "hbm", "kernel": "k_lsmc_persistent<3, ring>: 50 exercise dates in one launch "
                                      "(induction only)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic,
            "traffic_note": "DRAM bytes per launch measured by ncu (profiles/r02_ncu_lsmc_summary.md): "
                            "the 1.68 GB path store read once + cashflow write-backs; `achieved` is "
                            "these bytes over the launch time.  The 32 B per path-date of SURVEY 8 d4 "
                            "(6.7 GB) are mostly L2 hits at this size, so the launch is latency / "
                            "issue bound, not HBM bound",
            "algorithmic_gbs": (1 << 22) * 50 * 32 / (best * 1e-3) / 1e9,
            "per_unit": "32 B per path-date algorithmic (SURVEY 8 d4); 8.8 B per path-date from DRAM",
            "floor_ms": floor_ms,
            "floor_note": "whole pricing: the store written once by the fill and read once by the "
                          "induction at the measured HBM bandwidth",
            "price_ms": price_ms, "frac_of_floor": floor_ms / price_ms,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
            "kernel_ms": best}


def parity_multi(cx, w, dist_moments_seed=2000):
    """N > 1: the partitioned job must reproduce the single-GPU job (rank 0 re-runs it whole)."""
    import numpy as np

    import optpricing_b200 as ob
    from optpricing_b200 import _engine
    eng, torch = cx.eng, cx.torch
    out = {}
    # C2 moments: strong job, same seed, all ranks vs rank 0 alone
    n_draws = w["n_paths"] // 2
    mdl = _engine.make_model(w["kind"], w["params"], w["r"], w["q"])
    off, cnt = ob.techniques.monte_carlo.shard_draws(n_draws, cx.rank, cx.world)
    mom = torch.zeros(_engine.NMOM, dtype=torch.float64, device=cx.dev)
    sim = eng.make_sim(w["S0"], w["T"], n_draws, w["n_steps"], True, 77, _engine.CORR_REFERENCE_EUROPEAN,
                       off, cnt)
    eng.european_async(mdl, sim, [w["K"]], [1], mom)
    eng.group_all_reduce()(mom)
    got = mom.cpu().numpy()
    if cx.rank == 0:
        sim1 = eng.make_sim(w["S0"], w["T"], n_draws, w["n_steps"], True, 77,
                            _engine.CORR_REFERENCE_EUROPEAN)
        eng.european_async(mdl, sim1, [w["K"]], [1], mom)
        one = mom.cpu().numpy()
        out["c2_moments_max_rel_diff"] = float(np.max(np.abs(got - one) / np.abs(one)))
    cx.barrier()
    # C4 price: NVLink in-kernel exchange and NCCL per-date exchange vs one GPU
    aput = ob.Option(strike=110.0, maturity=1.0, option_type=ob.OptionType.PUT)
    stock, rate, model = ob.Stock(spot=100.0, dividend=0.01), ob.Rate(rate=0.05), ob.BSMModel({"sigma": 0.2})
    prices = {}
    for ex in ("nvlink", "nccl"):
        t = ob.AmericanMonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=42, lsm_degree=3,
                                           exchange=ex)
        prices[ex] = t.price(aput, stock, model, rate).price
        cx.barrier()
    if cx.rank == 0:
        t = ob.AmericanMonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=42, lsm_degree=3,
                                           distributed=False)
        one = t.price(aput, stock, model, rate).price
        out["c4_price_single_gpu"] = one
        for ex in prices:
            out[f"c4_price_{ex}_rel_diff"] = abs(prices[ex] - one) / abs(one)
    cx.barrier()
    return out


def run_ours(args, w):
    import ctypes

    import optpricing_b200 as ob
    from optpricing_b200 import _engine
    cx = Ctx(args)
    is_c2 = args.workload == "c2"
    sampler = ClockSampler(cx.local) if cx.rank == 0 else None   # started early: nvidia-smi is slow
    main = c2_device(cx, args, w, args.scaling, args.steps, sampler)
    e2e = c2_e2e(cx, args, w, args.scaling, args.steps)
    stats_dev = ob.techniques.monte_carlo.summarize(
        main["moments"], w["r"], w["T"],
        ob.techniques.monte_carlo.control_mean(w["kind"], w["params"], w["S0"], w["r"], w["q"],
                                               w["T"], w["n_steps"]))
    other = None
    if cx.world > 1:
        oscal = "weak" if args.scaling == "strong" else "strong"
        od = c2_device(cx, args, w, oscal, max(3, args.steps // 2))
        oe = c2_e2e(cx, args, w, oscal, max(3, args.steps // 2))
        other = {"scaling": oscal, "paths_per_step": od["n_job"], "value": od["value"],
                 "ms_per_step": od["ms_per_step"], "kernel_ms": od["kernel_ms"],
                 "e2e_value": oe["value"]}
    sec, par = {}, None
    if is_c2 and not args.no_secondary:
        for name in ("c1", "c3", "c4", "c5"):
            sec[name] = secondary(cx, name, 5 if name != "c1" else 20)
        if cx.world > 1:
            par = parity_multi(cx, w)
    if cx.rank != 0:
        cx.close()
        return
    roofline = fp64_roofline(cx, w, main["cnt"], main["kernel_ms"], main["clocks"], cx.world, is_c2)
    cpu = cpu_baseline_c2(w)
    h2d = ctypes.sizeof(_engine.OptmcModel) + ctypes.sizeof(_engine.OptmcSim) + 8 + 4
    line = {
        "metric": "path-steps/sec (fp64)", "value": main["value"], "unit": "path-steps/s",
        "n_gpus": cx.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "paths_per_step": main["n_job"], "n_steps": w["n_steps"],
                   "partition": f"draws sharded over {cx.world} rank(s), one all-reduce of 6 doubles"
                                + (" by k_peer_allreduce over NVLink peer memory (NCCL sets up the group)"
                                   if cx.world > 1 else ""),
                   "l2": "256 MiB written between timed steps (outside the per-step events); "
                         "the kernel reads no HBM input"},
        "e2e": {"value": e2e["value"], "unit": "path-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8 * _engine.NMOM, "api": "MonteCarloTechnique.price",
                "price": e2e["price"], "stderr": e2e["stderr"]},
        "gpu_launches": main["launches"], "clocks": main["clocks"], "roofline": roofline,
        "cpu_baseline": cpu,
        "check": {"price": stats_dev["price"], "stderr": stats_dev["stderr"],
                  "control_mean": stats_dev["control_mean"],
                  "control_expected": stats_dev["control_expected"],
                  "wall_s_timed_region": main["wall"]},
    }
    if other:
        line[other["scaling"]] = other
    if sec:
        line["secondary"] = sec
    if par:
        line["parity_multi"] = par
    print(json.dumps(line), flush=True)
    cx.close()


def run_secondary_only(args):
    cx = Ctx(args)
    res = secondary(cx, args.workload, args.steps)
    if cx.rank == 0:
        line = {"metric": "path-steps/sec (fp64)", "value": res["value"], "unit": "path-steps/s",
                "n_gpus": cx.world, "steps": args.steps, "warmup": 3,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": res["workload"],
                           "l2": "path store / terminal buffers exceed L2 or no HBM input; no flush "
                                 "between steps"},
                "e2e": {**res["e2e"], "h2d_bytes_per_step": 212, "d2h_bytes_per_step": 48},
                "gpu_launches": res["gpu_launches"],
                **{k: v for k, v in res.items() if k in ("roofline", "cpu_baseline", "results",
                                                         "result_sample", "simulations_per_step")}}
        print(json.dumps(line), flush=True)
    cx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c3", "c4", "c5"])
    ap.add_argument("--scaling", default="strong", choices=("weak", "strong"))
    ap.add_argument("--no-secondary", action="store_true",
                    help="C2 line only (skip the C1/C3/C4/C5 block and the multi-GPU parity block)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if args.workload not in WORKLOADS:
            raise SystemExit("--impl reference is defined for the c1/c2 workloads")
        return run_reference(args, WORKLOADS[args.workload])
    if args.workload in ("c3", "c4", "c5"):
        return run_secondary_only(args)
    run_ours(args, WORKLOADS[args.workload])


if __name__ == "__main__":
    main()
