"""Development probe: the PF-grid Metropolis kernel on C4-shaped tables (not a bench, not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t = time.time()
tabs = synthetic.make_pf_tables(n_states=S)
print("tables %.1fs, Nc %.0f MB Nh %.0f MB" % (time.time() - t, tabs["restraints"][0]["pf"]["Nc"].nbytes / 1e6,
                                               tabs["restraints"][0]["pf"]["Nh"].nbytes / 1e6), flush=True)
T = engine.DeviceTables(tabs)
for n_chains, n_steps in [(4096, 2000), (32768, 1000), (131072, 500)]:
    blk = T.chains(n_chains, seed=1)
    blk.launch(100); blk.sync()
    best = 1e9
    for rep in range(3):
        blk.launch(n_steps); blk.sync()
        best = min(best, blk.last_kernel_ms()[0])
    o = blk.fetch()
    rate = n_chains * n_steps / (best * 1e-3)
    print("chains %7d steps %5d  %.3f ms  %.3e chain-steps/s  (%.0f GB/s of <N> rows)  acc %.1f%% sep %s" % (
        n_chains, n_steps, best, rate, rate * 16 * 107 / 1e9, 100.0 * o["accepted"].mean() / o["total"][0],
        np.round(o["sep_accepted"].mean(0) / o["total"][0], 3)), flush=True)
    blk.close()
