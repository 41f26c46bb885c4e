"""Development probe: where the host-buffer (e2e) path spends its time.  Not a bench, not a test."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from biceps_b200 import _cabi, engine, precompute, synthetic

ens = synthetic.make_ensemble("C3", seed=0, n_lambda=8)
t0 = time.perf_counter()
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], device=0)
print("precompute.build_tables %.1f ms" % ((time.perf_counter() - t0) * 1e3))
packed = _cabi.PackedTables(tabs, pinned=True)      # inputs in page-locked host memory
n = 4096
lam = (np.arange(n) // (n // 8)).astype(np.int32)
o = None
for rep in range(4):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    T2 = engine.DeviceTables(packed, device=0); t.append(time.perf_counter())
    b2 = T2.chains(n, 7, chain_lambda=lam); t.append(time.perf_counter())
    o = b2.sample(100000, burn=0, traj_every=100, pinned=True, out=o); t.append(time.perf_counter())
    km = b2.last_kernel_ms()[0]
    b2.close(); T2.close(); t.append(time.perf_counter())
    names = ["create", "chains", "sample(kernel+D2H)", "close"]
    print("rep %d: " % rep + "  ".join("%s %.2f ms" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t[:-1], t[1:])) +
          "  total %.2f ms  kernel %.2f ms" % ((t[-1] - t[0]) * 1e3, km))
