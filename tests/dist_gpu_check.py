#!/usr/bin/env python
"""
Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tests/dist_gpu_check.py

Every rank prices the same contracts (a) through the distributed techniques
(draws sharded by rank, moments / power sums all-reduced over NVLink) and
(b) alone with distributed=False.  Streams are keyed by the global draw id, so
(a) must equal (b) up to summation order.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import optpricing_b200 as ob

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
call = ob.Option(100.0, 1.0, ob.OptionType.CALL)
aput = ob.Option(110.0, 1.0, ob.OptionType.PUT)
stock, rate = ob.Stock(100.0, dividend=0.01), ob.Rate(0.05)
fails = []
for name, model in (("heston", ob.HestonModel()), ("bates", ob.BatesModel()),
                    ("kou", ob.KouModel()), ("sabr_jump", ob.SABRJumpModel())):
    kw = dict(n_paths=(1 << 20) + 6, n_steps=32, seed=77)
    d = ob.MonteCarloTechnique(**kw).price(call, stock, model, rate).price
    s = ob.MonteCarloTechnique(distributed=False, **kw).price(call, stock, model, rate).price
    ok = abs(d - s) <= 1e-11 * abs(s)
    if rank == 0:
        print(f"european {name:9s} world={world}: distributed {d:.12f} single {s:.12f} {'ok' if ok else 'MISMATCH'}")
    if not ok:
        fails.append(name)
for name, model, deg in (("bsm", ob.BSMModel(), 3), ("heston", ob.HestonModel(), 2)):
    kw = dict(n_paths=(1 << 19) + 2, n_steps=20, seed=5, lsm_degree=deg)
    d = ob.AmericanMonteCarloTechnique(**kw).price(aput, stock, model, rate).price
    s = ob.AmericanMonteCarloTechnique(distributed=False, **kw).price(aput, stock, model, rate).price
    ok = abs(d - s) <= 1e-9 * abs(s)
    if rank == 0:
        print(f"american {name:9s} world={world}: distributed {d:.12f} single {s:.12f} {'ok' if ok else 'MISMATCH'}")
    if not ok:
        fails.append("lsmc_" + name)
t = torch.tensor([len(fails)], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
if rank == 0:
    print("dist_gpu_check:", "PASS" if t.item() == 0 else f"FAIL {fails}")
sys.exit(0 if t.item() == 0 else 1)
