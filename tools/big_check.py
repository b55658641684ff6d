import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import optpricing_b200 as ob
call = ob.Option(100.0, 1.0, ob.OptionType.CALL)
aput = ob.Option(110.0, 1.0, ob.OptionType.PUT)
st, stq, rt = ob.Stock(100.0), ob.Stock(100.0, dividend=0.01), ob.Rate(0.05)
t = ob.MonteCarloTechnique(n_paths=1 << 31, n_steps=12, seed=1)
t0 = time.perf_counter(); p = t.price(call, st, ob.BSMModel(), rt).price; dt = time.perf_counter() - t0
print(f"BSM 2^31 paths x 12 steps: {p:.6f} +- {t.last_stats['stderr']:.6f} (closed form 10.450584) in {dt*1e3:.1f} ms, n_samples {t.last_stats['n_samples']:.0f}")
t = ob.MonteCarloTechnique(n_paths=(1 << 32) + 2, n_steps=2, seed=2)
t0 = time.perf_counter(); p = t.price(call, st, ob.HestonModel(), rt).price; dt = time.perf_counter() - t0
print(f"Heston 2^32+2 paths x 2 steps: {p:.6f} +- {t.last_stats['stderr']:.6f} in {dt*1e3:.1f} ms, n_samples {t.last_stats['n_samples']:.0f}")
for n in (1 << 26, (1 << 27) + 6):
    a = ob.AmericanMonteCarloTechnique(n_paths=n, n_steps=50, seed=3, lsm_degree=3)
    a.price(aput, stq, ob.BSMModel(), rt)
    torch.cuda.synchronize(); t0 = time.perf_counter(); p = a.price(aput, stq, ob.BSMModel(), rt).price; torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"LSMC {n} paths x 50 dates (store {n*50*8/2**30:.1f} GiB): {p:.5f} (CRR 12.2777) in {dt*1e3:.1f} ms = {n*50/dt/1e9:.1f} G path-steps/s; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
