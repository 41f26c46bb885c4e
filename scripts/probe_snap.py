"""Development probe: cost of trajectory snapshots in the Metropolis kernel (C3, 4096 chains)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, precompute, synthetic
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs)
lam = (np.arange(4096) % 8).astype(np.int32)
blk = T.chains(4096, 1, chain_lambda=lam)
for te, burn in [(0, 0), (100, 0), (100, 2), (7, 0), (1000, 0)]:
    blk.launch(2000); blk.sync()
    best = 1e9
    for _ in range(3):
        blk.launch(20000, burn=burn, traj_every=te); blk.sync()
        best = min(best, blk.last_kernel_ms()[0])
    print("traj_every %4d burn %d: %.3f ms  %.3e chain-steps/s" % (te, burn, best, 4096 * (20000 + burn) / (best * 1e-3)))
