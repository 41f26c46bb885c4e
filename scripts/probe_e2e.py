"""Where an end-to-end step (bench.py's e2e_step) spends its time: table upload, chain block, sampling, close."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from biceps_b200 import engine, precompute, synthetic, _cabi

n, n_mcmc = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], device=0)
packed = _cabi.PackedTables(tabs, pinned=True)
lam = (np.arange(n) // (n // 8)).astype(np.int32)
o = None
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    T2 = engine.DeviceTables(packed, device=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b2 = T2.chains(n, 7, chain_id0=0, chain_lambda=lam)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    o = b2.sample(n_mcmc, burn=0, traj_every=100, pinned=True, out=o)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    kms = b2.last_kernel_ms()
    b2.close(); T2.close()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("rep %d: tables %.2f ms, chains %.2f ms, sample %.2f ms (kernels %.2f ms in %d launches, %s), close %.2f ms, total %.2f ms"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), kms[0], kms[1], b2.last_kernel() if False else "-", 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
