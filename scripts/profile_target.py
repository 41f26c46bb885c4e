"""Short single-purpose run for ncu: C3 tables (product precompute), then a few launches of the
Metropolis kernel.  usage: profile_target.py [n_chains] [n_steps] [reps] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, precompute, synthetic

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = sys.argv[4] if len(sys.argv) > 4 else "C3"
ens = synthetic.make_ensemble(cfg, seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs)
lam = (np.arange(n_chains) % 8).astype(np.int32)
blk = T.chains(n_chains, 1, chain_lambda=lam)
for _ in range(reps):
    blk.launch(n_steps)
    blk.sync()
    print("kernel", blk.last_kernel(), "ms", blk.last_kernel_ms())
