"""Development probe: saturated-regime timing of the thread-per-chain kernel, chains sorted by lambda or interleaved."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic, precompute
K = 8
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=K)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs, device=0)
for n_chains, n_steps in [(4096, 20000), (148 * 2048, 4000)]:
    for name, lam in [("interleaved", np.arange(n_chains) % K), ("sorted", np.arange(n_chains) // (n_chains // K + 1))]:
        blk = T.chains(n_chains, seed=1, chain_lambda=lam.astype(np.int32))
        blk.launch(200); blk.sync()
        best = 1e9
        for rep in range(3):
            blk.launch(n_steps); blk.sync()
            best = min(best, blk.last_kernel_ms()[0])
        print("kernel=%s chains %8d %-12s %.3f ms  %.3e chain-steps/s" % (os.environ.get("BCP_KERNEL", "auto"), n_chains, name, best, n_chains * n_steps / (best * 1e-3)), flush=True)
        blk.close()
