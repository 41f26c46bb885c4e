"""ncu target: every streaming stage kernel once warm at its named size -- K1/K2/K4 precompute at C3 size
(10^4 states), logZ, K6 u_kn and the two K7 MBAR passes + populations at C5 size (K = 11, N = 7.04e6), and the
K9 JSD bootstrap.  Each stage runs twice (first launch = warm-up); filter with ncu -k regex:..."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from biceps_b200 import _cabi as A, engine, precompute, synthetic

lib = A.lib()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream


def ptr(t):
    return C.c_void_p(t.data_ptr())


for S, J, kind, ref, G in [(10000, 300, A.BCP_KIND_SCALAR, 0, 1), (10000, 400, A.BCP_KIND_GAMMA, 1, 163),
                           (10000, 1000, A.BCP_KIND_SCALAR, 2, 1)]:
    e = rng.uniform(2, 6, J)
    model = (torch.from_numpy(e).to(dev)[None, :] * (1 + 0.15 * torch.randn((S, J), dtype=torch.float64, device=dev))).clamp_(min=0.5)
    exp = torch.from_numpy(e).to(dev); w = torch.ones(J, dtype=torch.float64, device=dev)
    gam = torch.from_numpy(precompute.log_grid(0.2, 5.0, 1.02)).to(dev)
    sse = torch.empty((S, G), dtype=torch.float64, device=dev)
    refsum = torch.empty(S, dtype=torch.float64, device=dev); p0 = torch.empty(J, dtype=torch.float64, device=dev); p1 = torch.empty(J, dtype=torch.float64, device=dev)
    obs = A.bcp_obs(kind, J, ref, 0, 1, G, C.cast(model.data_ptr(), A.c_double_p), C.cast(exp.data_ptr(), A.c_double_p),
                    C.cast(w.data_ptr(), A.c_double_p), C.cast(gam.data_ptr(), A.c_double_p))
    for _ in range(2):
        A.check(lib.bcp_precompute_dev(C.byref(obs), S, 0, C.c_void_p(st), ptr(sse), ptr(refsum), ptr(p0), ptr(p1)), "precompute")
    torch.cuda.synchronize()

ens = synthetic.make_ensemble("C3", seed=0, n_lambda=11)
tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])      # logz_kernel runs in here
T = engine.DeviceTables(tabs)
S, K, per = 10000, 11, 64 * 10000
N = K * per
st_h = rng.integers(0, S, N).astype(np.int32)
idx_h = np.stack([rng.integers(gl // 2 - 3, gl // 2 + 3, N) for gl in T.grid_len], axis=1).astype(np.int32)
d_state = torch.from_numpy(st_h).to(dev); d_idx = torch.from_numpy(idx_h).to(dev)
d_u = torch.empty((K, N), dtype=torch.float64, device=dev)
for _ in range(2):
    A.check(lib.bcp_ukn_dev(T._h, N, ptr(d_state), ptr(d_idx), ptr(d_u), C.c_void_p(st)), "ukn")
torch.cuda.synchronize()

en = torch.from_numpy(tabs["energy"]).to(dev)
samp_state = torch.cat([torch.multinomial(torch.softmax(-en[k], 0), per, replacement=True) for k in range(K)])
base = torch.from_numpy(rng.normal(50.0, 3.0, N)).to(dev)
d_u2 = (en[:, samp_state] + base[None, :]).contiguous()
N_k = np.full(K, per, dtype=np.int64)
f = np.zeros(K); s_out = np.zeros(K); M_out = np.zeros((K, K))
for hess in (0, 1, 0, 1):
    A.check(lib.bcp_mbar_pass_dev(ptr(d_u2), N, K, A.i64ptr(N_k), A.dptr(f), hess, 0, C.c_void_p(st), A.dptr(s_out), A.dptr(M_out)), "pass")
fk = np.zeros(K); theta = np.zeros((K, K)); it = C.c_int32(0)
d_st32 = samp_state.to(torch.int32)
pops = np.zeros((S, K)); dpops = np.zeros((S, K))
A.check(lib.bcp_mbar_dev(ptr(d_u2), A.i64ptr(N_k), K, N, ptr(d_st32), S, 0, C.c_void_p(st), 0.0, 0, A.dptr(fk), A.dptr(theta),
                         A.dptr(pops), A.dptr(dpops), C.byref(it)), "mbar")
torch.cuda.synchronize()
print("stages ok; mbar iterations", it.value)
