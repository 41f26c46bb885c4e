"""What separates the kernels of an end-to-end step from the steady-state launch: fresh against warm chains, with and without
snapshots (kernel milliseconds of one sample call of 100 000 steps, 4096 chains, C3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from biceps_b200 import engine, precompute, synthetic, _cabi

n, n_mcmc = 4096, 100000
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], device=0)
T = engine.DeviceTables(_cabi.PackedTables(tabs, pinned=True), device=0)
lam = (np.arange(n) // (n // 8)).astype(np.int32)
for te in (100, 0, 1000):
    for rep in range(2):
        b = T.chains(n, 7 + rep, chain_id0=0, chain_lambda=lam)
        out = []
        for call in range(3):
            b.sample(n_mcmc, burn=0, traj_every=te, pinned=True)
            out.append("%.2f ms / %d launches" % b.last_kernel_ms())
        b.close()
        print("traj_every %4d: fresh %s | second call %s | third call %s" % (te, *out))
