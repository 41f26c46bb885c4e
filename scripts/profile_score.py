"""Development probe: the C5 score pipeline (sample -> bcp_chains_score) with wall-clock phase times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, precompute, synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ens = synthetic.make_ensemble("C3", seed=0, n_lambda=11)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs)
n = 11 * 64
lam = (np.arange(n) % 11).astype(np.int32)
for rep in range(3):
    b = T.chains(n, 5 + rep, chain_lambda=lam)
    t0 = time.perf_counter()
    b.launch(steps, traj_every=100); b.sync()
    t1 = time.perf_counter()
    r = b.score(populations=True)
    t2 = time.perf_counter()
    r2 = b.score(populations=False)
    t3 = time.perf_counter()
    print("rep %d sample %.1f ms  score+pops %.1f ms  score only %.1f ms  iters %d  f[-1] %.6f +- %.6f" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, r["iters"], r["f"][-1], r["dDelta_f"][0, -1]))
    b.close()
