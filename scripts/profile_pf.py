"""Short single-purpose run for ncu: C4 PF tables, a few launches of the PF Metropolis kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic
n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
tabs = synthetic.make_pf_tables(n_states=100000)
T = engine.DeviceTables(tabs)
blk = T.chains(n_chains, 1)
for _ in range(3):
    blk.launch(n_steps); blk.sync()
    print("kernel ms", blk.last_kernel_ms())
