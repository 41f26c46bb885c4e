"""Development probe: the reference-facing Python API (PosteriorSampler.sample) on the cineromycin-shaped
synthetic ensemble: one chain x 1e6 steps (BASELINE configs[0]) and 4096 chains x 1e5 steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import biceps_b200 as biceps
from biceps_b200 import synthetic
ens = synthetic.make_ensemble("C1s", seed=3)
for nchains, nsteps in ((1, 1000000), (4096, 100000)):
    rs = []
    for r in ens["restraints"]:
        noe = r["name"] == "noe"
        opt = dict(ref=r["ref"], sigma=r["sigma"])
        if noe:
            opt["gamma"] = r["gamma"]
        rs.append(dict(cls="noe" if noe else "J", model=r["model"], exp=r["exp"], weight=r["weight"], options=opt))
    E = biceps.Ensemble.from_arrays(1.0, ens["energies"], rs)
    s = biceps.PosteriorSampler(E, nchains=nchains, seed=1)
    s.sample(1000)
    t0 = time.perf_counter()
    s.sample(nsteps)
    dt = time.perf_counter() - t0
    print("nchains %5d nsteps %8d: %.3f s  %.3e chain-steps/s  (%d snapshots in traj)" % (nchains, nsteps, dt, nchains * nsteps / dt, len(s.traj.trajectory)))
