"""Development probe: the speculative kernel on C3 tables for a given number of chains (argv[1])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic
from oracle import precompute_oracle as PO
K = 8
n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=K)
tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
T = engine.DeviceTables(tabs, device=0)
lam = (np.arange(n_chains) % K).astype(np.int32)
blk = T.chains(n_chains, seed=1, chain_lambda=lam)
blk.launch(2000); blk.sync()
best = 1e9
for rep in range(3):
    blk.launch(20000); blk.sync()
    best = min(best, blk.last_kernel_ms()[0])
print("L=%s chains %d: %.3f ms per 20000 steps = %.0f cycles/step at 1965 MHz" % (os.environ.get("BCP_SPEC_L", "4"), n_chains, best, best * 1e-3 * 1.965e9 / 20000))
