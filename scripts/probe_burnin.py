"""Burn-in profile of the table-driven kernel: consecutive short launches from a fresh ensemble, time and rolled-back
batches per launch, next to the speculative kernel.  usage: probe_burnin.py [n_chains] [steps_per_launch] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, precompute, synthetic

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
cfg = sys.argv[4] if len(sys.argv) > 4 else "C3"
ens = synthetic.make_ensemble(cfg, seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs)
lam = (np.arange(n_chains) % 8).astype(np.int32)
for kern in ("tab", "spec"):
    os.environ["BCP_KERNEL"] = kern
    blk = T.chains(n_chains, 1, chain_lambda=lam)
    ms = []
    for _ in range(reps):
        blk.launch(n_steps)
        blk.sync()
        ms.append(blk.last_kernel_ms()[0])
    print(kern, "ms per launch of", n_steps, "steps:", " ".join("%.3f" % m for m in ms))
