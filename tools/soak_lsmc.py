#!/usr/bin/env python
"""GPU soak: thousands of persistent-kernel Longstaff-Schwartz pricings back to back (single GPU,
or jointly over NVLink under torchrun); every repetition of a seed must give the identical price
and no exchange may time out."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = 0
if world > 1:
    import torch.distributed as dist
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
aput = ob.Option(110.0, 1.0, ob.OptionType.PUT)
stock, rate, model = ob.Stock(100.0, dividend=0.01), ob.Rate(0.05), ob.BSMModel()
t0 = time.perf_counter()
bad = 0
for n_paths, n_steps, reps in ((20_000, 20, 3000), (1 << 20, 50, 500), (1 << 22, 50, 100)):
    ref = None
    for i in range(reps):
        t = ob.AmericanMonteCarloTechnique(n_paths=n_paths, n_steps=n_steps, seed=5, lsm_degree=3)
        p = t.price(aput, stock, model, rate).price
        if ref is None:
            ref = p
        elif p != ref:
            bad += 1
    if rank == 0:
        print(f"world={world} {n_paths} x {n_steps}: {reps} pricings, price {ref:.10f}, mismatches so far {bad}", flush=True)
if rank == 0:
    print(f"soak done in {time.perf_counter() - t0:.1f} s: {'PASS' if bad == 0 else 'FAIL'}")
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if bad == 0 else 1)
