import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from biceps_b200 import engine
from helpers import cineromycin_tables
z, t0, t1 = cineromycin_tables()
T = engine.DeviceTables(t1, device=0)
np.random.seed(1)
blk = T.chains(1, 0, init_state=np.random.randint(0, T.n_states, 1).astype(np.int32), rng_mode=1)
blk.sample_numpy_stream(1000)
t = time.perf_counter(); blk.sample_numpy_stream(1000000, traj_every=100); dt = time.perf_counter() - t
print("MT mode: 1e6 steps in %.2f s (%.2e steps/s, kernel %.1f ms)" % (dt, 1e6 / dt, blk.last_kernel_ms()[0]))
b2 = T.chains(1, 3)
b2.sample(1000)
t = time.perf_counter(); b2.sample(1000000, traj_every=100); dt = time.perf_counter() - t
print("Philox, one chain: 1e6 steps in %.2f s (kernel %.1f ms)" % (dt, b2.last_kernel_ms()[0]))
