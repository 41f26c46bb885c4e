"""ncu target: a few MBAR passes (K = 11, N = 7.04e6) on a device-resident u_kn."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from biceps_b200 import _cabi as A
lib = A.lib()
K, N = 11, 11 * 64 * 10000
g = torch.Generator(device="cuda").manual_seed(0)
u = (50.0 + 3.0 * torch.randn((K, N), dtype=torch.float64, device="cuda", generator=g)).contiguous()
N_k = np.full(K, N // K, dtype=np.int64); f = np.zeros(K); s = np.zeros(K); M = np.zeros((K, K))
for hess in (0, 1, 0, 1):
    A.check(lib.bcp_mbar_pass_dev(C.c_void_p(u.data_ptr()), N, K, A.i64ptr(N_k), A.dptr(f), hess, 0, C.c_void_p(0), A.dptr(s), A.dptr(M)), "pass")
torch.cuda.synchronize()
print("s", s[:3])
