"""Development probe (not a bench, not a test): times the Metropolis kernel on C3-shaped tables for
several chain counts.  Tables come from the numpy oracle here, so this file is tooling only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic
from oracle import precompute_oracle as PO

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
K = 8
t = time.time()
ens = synthetic.make_ensemble(cfg, seed=0, n_lambda=K)
tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
print("tables built in %.1fs kernel=%s" % (time.time() - t, os.environ.get("BCP_KERNEL", "auto")), flush=True)
T = engine.DeviceTables(tabs, device=0)
for n_chains, n_steps in [(4096, 20000), (4096 * 8, 10000), (148 * 2048, 4000), (148 * 2048 * 2, 2000)]:
    lam = (np.arange(n_chains) % K).astype(np.int32)
    blk = T.chains(n_chains, seed=1, chain_lambda=lam)
    blk.launch(200); blk.sync()
    best = 1e9
    for rep in range(3):
        blk.launch(n_steps); blk.sync()
        ms, nl = blk.last_kernel_ms()
        best = min(best, ms)
    o = blk.fetch()
    print("chains %8d steps %6d  %.3f ms  %.3e chain-steps/s  acc %.1f%%" % (
        n_chains, n_steps, best, n_chains * n_steps / (best * 1e-3), 100.0 * o["accepted"].mean() / o["total"][0]), flush=True)
    blk.close()
