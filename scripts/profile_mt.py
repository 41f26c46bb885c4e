"""ncu target: the reference-stream kernel, one chain x 200 000 steps on the cineromycin tables."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from biceps_b200 import engine
from helpers import cineromycin_tables
z, t0, t1 = cineromycin_tables()
T = engine.DeviceTables(t1, device=0)
np.random.seed(1)
blk = T.chains(1, 0, init_state=np.random.randint(0, T.n_states, 1).astype(np.int32), rng_mode=1)
blk.sample_numpy_stream(1000)
blk.sample_numpy_stream(200000, traj_every=100)
print("ok")
