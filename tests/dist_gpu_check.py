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
# the engine's own all-reduce kernel (NVLink peer memory) against NCCL's, several sizes and calls
from optpricing_b200 import _engine
eng = _engine.get_engine()
ar = eng.group_all_reduce()
if rank == 0:
    print("group all-reduce:", "k_peer_allreduce over NVLink peer memory" if ar == eng.peer_all_reduce
          else "torch.distributed (NCCL)")
for n in (1, 6, 384, 1536, 4000):
    for rep in range(3):
        g = torch.Generator(device="cuda").manual_seed(1000 * rank + n + rep)
        a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        b = a.clone()
        ar(a)
        dist.all_reduce(b)
        if not torch.allclose(a, b, rtol=1e-14, atol=1e-14):
            fails.append(f"peer_allreduce_{n}")
import time
for name, model, deg in (("bsm", ob.BSMModel(), 3), ("heston", ob.HestonModel(), 2)):
    kw = dict(n_paths=(1 << 19) + 2, n_steps=20, seed=5, lsm_degree=deg)
    s = ob.AmericanMonteCarloTechnique(distributed=False, **kw).price(aput, stock, model, rate).price
    for exchange in ("nvlink", "nccl", "nvlink"):
        d = ob.AmericanMonteCarloTechnique(exchange=exchange, **kw).price(aput, stock, model, rate).price
        ok = abs(d - s) <= 1e-9 * abs(s)
        if rank == 0:
            print(f"american {name:9s} world={world} {exchange:6s}: distributed {d:.12f} single {s:.12f} "
                  f"{'ok' if ok else 'MISMATCH'}")
        if not ok:
            fails.append(f"lsmc_{name}_{exchange}")
# timing of the two exchanges at BASELINE config 4 (2^22 paths x 50 dates, cubic basis)
for exchange in (() if os.environ.get("DIST_CHECK_QUICK") else ("nvlink", "nccl")):
    t = ob.AmericanMonteCarloTechnique(n_paths=1 << 22, n_steps=50, seed=1, lsm_degree=3, exchange=exchange)
    for _ in range(3):
        p = t.price(aput, stock, ob.BSMModel(), rate).price
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(10):
        p = t.price(aput, stock, ob.BSMModel(), rate).price
    torch.cuda.synchronize(); dist.barrier(); ms = (time.perf_counter() - t0) * 100
    if rank == 0:
        print(f"C4 american put 2^22 x 50 deg 3, world={world}, exchange={exchange}: {ms:.3f} ms per price "
              f"({(1 << 22) * 50 / ms / 1e6:.1f} G path-steps/s), price {p:.4f}")
t = torch.tensor([len(fails)], device="cuda")
dist.all_reduce(t)
dist.destroy_process_group()
if rank == 0:
    print("dist_gpu_check:", "PASS" if t.item() == 0 else f"FAIL {fails}")
sys.exit(0 if t.item() == 0 else 1)
