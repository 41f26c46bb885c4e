import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic
from oracle import precompute_oracle as PO
K = 8
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=K)
tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
T = engine.DeviceTables(tabs, device=0)
for n_chains, n_steps in [(4096, 20000)]:
    lam = (np.arange(n_chains) % K).astype(np.int32)
    blk = T.chains(n_chains, seed=1, chain_lambda=lam)
    blk.launch(2000); blk.sync()
    best = 1e9
    for rep in range(3):
        blk.launch(n_steps); blk.sync()
        ms, nl = blk.last_kernel_ms()
        best = min(best, ms)
    o = blk.fetch()
    print("L=%s chains %8d steps %6d  %.3f ms  %.3e chain-steps/s  acc %.1f%% sumE %.6f" % (os.environ.get("BCP_SPEC_L","4"),
        n_chains, n_steps, best, n_chains * n_steps / (best * 1e-3), 100.0 * o["accepted"].mean() / o["total"][0], o["energy"].sum()), flush=True)
    blk.close()
