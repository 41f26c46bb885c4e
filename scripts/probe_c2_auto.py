"""Development probe: the automatic kernel choice on the C2 shape (a long first launch holds the trial launches; the verdict
sticks to the chain block)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from biceps_b200 import engine, synthetic, precompute
os.environ.pop("BCP_KERNEL", None)
ens = synthetic.make_ensemble("C2", seed=0, n_lambda=8)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
T = engine.DeviceTables(tabs, device=0)
n = 8192
blk = T.chains(n, seed=1, chain_lambda=(np.arange(n) % 8).astype(np.int32))
blk.launch(40000); blk.sync()
print("first launch: kernel", blk.last_kernel(), "ms", blk.last_kernel_ms())
for rep in range(3):
    blk.launch(20000); blk.sync()
    ms = blk.last_kernel_ms()[0]
    print("20000 steps: kernel %s %.3f ms %.3e chain-steps/s" % (blk.last_kernel(), ms, n * 20000 / (ms * 1e-3)))
