#!/usr/bin/env python
"""GPU experiment, run under torchrun: where one exercise date of the JOINT persistent
Longstaff-Schwartz launch (paths sharded over the ranks, Gram sums exchanged over NVLink
inside the kernel) spends its time.  Block 0 of every rank stamps clock64 at four points
per date; rank 0 prints the medians over the dates, and the device time of a whole price().

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29519 tools/lsmc_phases_dist.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import optpricing_b200 as ob
from optpricing_b200 import _engine

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
eng = _engine.get_engine()
n_steps = 50
aput = ob.Option(110.0, 1.0, ob.OptionType.PUT)
stock, rate = ob.Stock(100.0, dividend=0.01), ob.Rate(0.05)
for n_paths in (1 << 20, 1 << 22):
    t = ob.AmericanMonteCarloTechnique(n_paths=n_paths, n_steps=n_steps, seed=1, lsm_degree=3)
    tim = torch.zeros((n_steps + 1) * 4, dtype=torch.int64, device=eng.device)
    for rep in range(4):
        eng.set_ptr("lsmc_timing", tim if rep == 3 else None)
        p = t.price(aput, stock, ob.BSMModel(), rate).price
    eng.set_ptr("lsmc_timing", None)
    torch.cuda.synchronize()
    s = tim.cpu().numpy().reshape(n_steps + 1, 4).astype(np.float64)
    rows = s[1:n_steps][::-1]
    poll, solve, sweep = rows[:, 1] - rows[:, 0], rows[:, 2] - rows[:, 1], rows[:, 3] - rows[:, 2]
    publish = np.r_[rows[1:, 0], s[0, 0]] - rows[:, 3]
    med = lambda a: float(np.median(a)) / 1965.0
    # device time of whole pricings, max over ranks
    best = 1e9
    for rep in range(12):
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t.price(aput, stock, ob.BSMModel(), rate)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep >= 2:
            best = min(best, ms.item())
    for r in range(world):
        dist.barrier()
        if r == rank and (r == 0 or r == world - 1):
            print(f"world {world} rank {rank} n_paths {n_paths}: per date [us] local poll+sum {med(poll):.2f}  "
                  f"peer exchange+solve {med(solve):.2f}  sweep {med(sweep):.2f}  block_sum+publish {med(publish):.2f}  "
                  f"total {med(poll + solve + sweep + publish):.2f};  price() {best:.3f} ms (best of 10, max over ranks)  "
                  f"price {p:.6f}", flush=True)
dist.destroy_process_group()
