"""Development probe: kernel choice on the C2 shape (1k states, R=5, 8 lambdas x 1024 chains)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic, precompute
ens = synthetic.make_ensemble("C2", seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs, device=0)
for kern in ("tab", "ws", "spec", "coop", "fast"):
    os.environ["BCP_KERNEL"] = kern
    for n_chains, n_steps in [(1024, 20000), (8192, 20000), (65536, 4000)]:
        lam = (np.arange(n_chains) % 8).astype(np.int32)
        blk = T.chains(n_chains, seed=1, chain_lambda=lam)
        blk.launch(200); blk.sync()
        best = 1e9
        for rep in range(3):
            blk.launch(n_steps); blk.sync()
            best = min(best, blk.last_kernel_ms()[0])
        print("%-5s (ran %s) chains %6d  %.3f ms  %.3e chain-steps/s" % (kern, blk.last_kernel(), n_chains, best, n_chains * n_steps / (best * 1e-3)), flush=True)
        blk.close()
