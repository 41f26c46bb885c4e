"""ncu target: the autocorrelation kernel on 64 series x 1e5 samples x 1e4 lags (device-resident)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biceps_b200 import _cabi as A

lib = A.lib()
n, T, mt = 64, 100000, 10000
x = torch.randn((n, T), dtype=torch.float64, device="cuda") + 3.0
curves = torch.empty((n, mt + 1), dtype=torch.float64, device="cuda")
tau = torch.empty(n, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    A.check(lib.bcp_autocorr_dev(C.c_void_p(x.data_ptr()), n, T, mt, 1, 0, C.c_void_p(st), C.c_void_p(curves.data_ptr()),
                                 C.c_void_p(tau.data_ptr())), "autocorr")
torch.cuda.synchronize()
print("ok", float(curves[0, 0]))
