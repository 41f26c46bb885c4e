"""Small sampler runs for compute-sanitizer (memcheck / racecheck): every Metropolis kernel on a tiny ensemble
with frequent accepted state moves, snapshots and a state trace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic, mbar
from oracle import precompute_oracle as PO

ens = synthetic.make_ensemble("C1s", seed=3)
tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
T = engine.DeviceTables(tabs)
lam = (np.arange(40) % 2).astype(np.int32)
for kern in ("spec", "fast", "coop", "generic"):
    os.environ["BCP_KERNEL"] = kern
    blk = T.chains(40, 5, chain_lambda=lam)
    o = blk.sample(300, burn=10, traj_every=7, record_state_trace=True, trace_chains=3)
    r = blk.score()
    print(kern, "accepted", int(o["accepted"].sum()), "score", float(r["f"][-1]))
    blk.close()
ens2 = synthetic.make_ensemble("C2", seed=1, n_states=50)
tabs2 = PO.build_tables(ens2["energies"], ens2["lambdas"], ens2["restraints"], exact_pow=False)
T2 = engine.DeviceTables(tabs2)
os.environ["BCP_KERNEL"] = "spec"
o = T2.chains(24, 2, chain_lambda=(np.arange(24) % 8).astype(np.int32)).sample(200, traj_every=5)
print("C2 spec accepted", int(o["accepted"].sum()))
pf = synthetic.make_pf_tables(n_states=40, n_obs=107)
T3 = engine.DeviceTables(pf)
o = T3.chains(8, 2).sample(100, traj_every=10)
print("pf accepted", int(o["accepted"].sum()))
# reference-stream mode and the convergence kernels
os.environ.pop("BCP_KERNEL", None)
np.random.seed(4)
blk = T.chains(1, 0, init_state=np.random.randint(0, T.n_states, 1).astype(np.int32), rng_mode=1)
o = blk.sample_numpy_stream(400, burn=5, traj_every=9, record_state_trace=True)
print("mt accepted", int(o["accepted"].sum()))
from biceps_b200 import convergence as BC
x = np.random.default_rng(0).standard_normal((5, 3001)).cumsum(axis=1)
for mt in (10, 1500, 3100):
    c, t = BC.autocorrelation(x, mt)
print("autocorr", float(t[0]))
idx = np.random.default_rng(1).integers(0, 50, 4000).astype(np.int32)
sel = np.sort(np.random.default_rng(2).choice(3000, 1500, replace=False)).astype(np.int32)
print("jsd", BC.jsd_tasks(idx, [(0, 3000, 1500, 0, 0, 50, 0), (1000, 3000, 1500, 0, 0, 50, 1), (7, 3993, 1996, 0, 9, 50, 2)], sel, seed=3),
      BC.param_hist(idx, [(0, 4000, 0, 0, 0, 50, 0)], 64).sum())
