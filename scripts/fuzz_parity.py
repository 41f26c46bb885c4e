"""Randomised parity fuzz (GPU): random canonical ensembles (1-8 restraints, with / without a gamma restraint,
small and large grids, reference potentials), random chain counts / steps / burn / snapshot spacing / lambda
layouts, every Metropolis kernel against the C oracle, bit for bit.  usage: fuzz_parity.py [n_cases] [seed] [big]
(big: grids of at least 32 values, at most 5 restraints, longer runs -- the shapes the table-driven kernel takes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from biceps_b200 import engine
from oracle import cbind
from helpers import assert_outputs_equal

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
big = len(sys.argv) > 3 and sys.argv[3] == "big"
kernels = ("tab", "ws", "spec", "fast", "auto") if big else ("tab", "ws", "spec", "fast", "coop", "generic", "auto")
bad = 0
for case in range(n_cases):
    R = int(rng.integers(1, 6 if big else 9))
    S = int(rng.choice([1, 2, 3, 7, 40, 150, 2000] if big else [1, 2, 3, 7, 40, 150]))
    K = int(rng.integers(1, 4))
    hg = bool(rng.integers(0, 2))
    flat = bool(rng.integers(0, 2))                  # flat tables -> frequent accepted state moves / wraps
    rest = []
    for q in range(R):
        nsig = int(rng.choice([32, 33, 40, 97]) if big else rng.choice([1, 2, 3, 9, 40]))
        sig = np.linspace(0.5, 1.5, nsig) if nsig > 1 else np.array([0.9])
        ndof = float(rng.uniform(1, 6))
        d = dict(kind="scalar", name="r%d" % q, ref="uniform", ndof=ndof, sigma=sig, nlsig=ndof * np.log(sig),
                 two_sig2=2.0 * sig ** 2, cnorm=float(rng.uniform(0, 3)), grids=[], refsum=None)
        scale = 1e-3 if flat else 1.0
        if hg and q == R - 1:
            G = int(rng.choice([32, 35, 60]) if big else rng.choice([2, 3, 5, 17, 60]))
            d.update(kind="gamma", grids=[np.linspace(0.8, 1.2, G)], sse=2.0 + scale * rng.uniform(0, 3, (S, G)))
        else:
            d["sse"] = 2.0 + scale * rng.uniform(0, 3, S)
        if rng.integers(0, 2):
            d["refsum"] = scale * rng.uniform(0, 1, S); d["ref"] = "exponential"
        rest.append(d)
    tabs = dict(n_states=S, energy=scale * rng.uniform(0, 2, (K, S)), logZ=rng.uniform(0, 1, K), restraints=rest)
    n = int(rng.choice([1, 5, 33, 70]))
    steps = int(rng.choice([1, 8, 61, 700, 4000] if big else [1, 3, 8, 61, 700]))
    burn = int(rng.choice([0, 1, 5, 40]))
    te = int(rng.choice([0, 7, 32, 41, 100] if big else [0, 1, 4, 7, 100]))
    trace = bool(rng.integers(0, 2))
    lam = rng.integers(0, K, n).astype(np.int32)
    kw = dict(burn=burn, traj_every=te, record_state_trace=trace)
    want = cbind.sample(tabs, n, 7 + case, steps, chain_lambda=lam, n_threads=4, **kw)
    want2 = cbind.sample(tabs, n, 7 + case, 50, chain_lambda=lam, n_threads=4, step0=steps + burn, resume=want, traj_every=3)
    T = engine.DeviceTables(tabs)
    ref3 = None
    if os.environ.get("FUZZ_ONLY") and int(os.environ["FUZZ_ONLY"]) != case:
        T.close()
        continue
    if os.environ.get("FUZZ_ONLY"):
        print("case", case, "grids", [(len(r_["sigma"]), len(r_["grids"][0]) if r_["grids"] else 0) for r_ in rest], flush=True)
    for kern in kernels:
        if kern == "auto":
            os.environ.pop("BCP_KERNEL", None)
        else:
            os.environ["BCP_KERNEL"] = kern
        try:
            blk = T.chains(n, 7 + case, chain_lambda=lam)
            got = blk.sample(steps, **kw)
            assert_outputs_equal(got, want)
            got2 = blk.sample(50, traj_every=3)               # second call continues the chains
            if big:                                           # ... and a third one on the kernel under test again
                got3 = blk.sample(200, traj_every=40)
                if ref3 is None:
                    ref3 = got3
                else:
                    assert_outputs_equal(got3, ref3)
            assert_outputs_equal(got2, want2)
            blk.close()
        except AssertionError as e:
            bad += 1
            if os.environ.get("FUZZ_ONLY"):
                for nm, (g_, w_) in (("call1", (got, want)), ("call2", (locals().get("got2"), want2))):
                    if g_ is None:
                        continue
                    for k in w_:
                        if not np.array_equal(g_[k], w_[k]):
                            d = np.argwhere(np.asarray(g_[k]) != np.asarray(w_[k]))
                            print("  ", kern, nm, k, "differs at", d[:6].tolist(), "got", np.asarray(g_[k])[tuple(d[0])], "want", np.asarray(w_[k])[tuple(d[0])], "n_diff", len(d))
            print("MISMATCH case %d kernel %s R=%d S=%d K=%d hg=%d flat=%d n=%d steps=%d burn=%d te=%d trace=%d: %s" % (
                case, kern, R, S, K, hg, flat, n, steps, burn, te, trace, str(e)[:80]), flush=True)
    T.close()
print("fuzz: %d cases x %d kernel choices, %d mismatches" % (n_cases, len(kernels), bad))
sys.exit(1 if bad else 0)
