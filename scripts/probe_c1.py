"""Development probe: the cineromycin-shaped ensemble (100 states, J + NOE, L = 2 lanes per chain)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic, precompute
ens = synthetic.make_ensemble("C1s", seed=3)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs)
for kern in ("spec", "fast"):
    os.environ["BCP_KERNEL"] = kern
    for n_chains, n_steps in [(1, 200000), (64, 100000), (4096, 20000), (16384, 10000)]:
        lam = (np.arange(n_chains) % 2).astype(np.int32)
        blk = T.chains(n_chains, 1, chain_lambda=lam)
        blk.launch(200); blk.sync()
        best = 1e9
        for _ in range(3):
            blk.launch(n_steps); blk.sync()
            best = min(best, blk.last_kernel_ms()[0])
        o = blk.fetch()
        print("%-5s chains %6d  %.3f ms  %.3e chain-steps/s  %.0f ns/step/chain-serial  state acc %.3f" % (
            kern, n_chains, best, n_chains * n_steps / (best * 1e-3), best * 1e6 / n_steps, o["sep_accepted"][:, -1].mean() / o["total"][0]), flush=True)
        blk.close()
