"""Stage benchmark of the convergence kernels (K8 autocorrelation, K9 JSD bootstrap):
device-resident CUDA-event times against the fp64 FMA peak, plus the end-to-end Convergence calls (host
buffers) next to the numpy restatement of the reference's loops on one host core (bounded sample).
Prints one JSON object."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from biceps_b200 import _cabi as A, convergence as BC
from oracle import convergence_oracle as CO      # the checker, timed as the CPU baseline only

lib = A.lib()
dev = torch.device("cuda:0")
prop = torch.cuda.get_device_properties(0)
# fp64 FMA peak: 64 DFMA / clk / SM (sm_100), at the clock nvidia-smi reports as max
sm_mhz = float(os.popen("nvidia-smi --query-gpu=clocks.max.sm --format=csv,noheader,nounits").read().split()[0])
peak_fma = prop.multi_processor_count * 64 * sm_mhz * 1e6
out = {"fp64_fma_peak_per_s": peak_fma, "sm_max_mhz": sm_mhz}


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def ar1(n, T, phi, seed=0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    e = torch.randn((n, T), dtype=torch.float64, device=dev, generator=g)
    # AR(1) by a scan in log space is overkill here: a moving average of white noise has the same cost profile
    k = torch.tensor([phi ** i for i in range(64)], dtype=torch.float64, device=dev)
    return torch.nn.functional.conv1d(e[:, None, :], k[None, None, :], padding=63)[:, 0, :T].contiguous() + 3.0


st = torch.cuda.current_stream().cuda_stream
for name, n, T, mt in [("K8_autocorr_P6_T1e4_tau1e4", 6, 10000, 10000), ("K8_autocorr_64x_T1e5_tau1e4", 64, 100000, 10000),
                       ("K8_autocorr_4096x6_T1e4_tau2000", 4096 * 6, 10000, 2000)]:
    x = ar1(n, T, 0.95)
    curves = torch.empty((n, mt + 1), dtype=torch.float64, device=dev)
    tau = torch.empty(n, dtype=torch.float64, device=dev)

    def run():
        A.check(lib.bcp_autocorr_dev(C.c_void_p(x.data_ptr()), n, T, mt, 1, 0, C.c_void_p(st), C.c_void_p(curves.data_ptr()),
                                     C.c_void_p(tau.data_ptr())), "autocorr")
    run(); torch.cuda.synchronize()
    ms = timed(run)
    # useful multiply-adds: lag tau overlaps T - 1 - tau samples
    fma = float(n) * sum(max(T - 1 - t, 0) for t in range(mt + 1))
    out[name] = {"ms": ms, "useful_fma": fma, "fma_per_s": fma / ms * 1e3, "frac_of_fp64_peak": fma / ms * 1e3 / peak_fma}
    if n <= 64:
        xs = x[:1].cpu().numpy()
        lags = min(mt, 400)
        t0 = time.perf_counter(); CO.g(xs[0], lags, True); dt = time.perf_counter() - t0
        cpu_fma = sum(max(T - 1 - t, 0) for t in range(lags + 1))
        out[name]["cpu_numpy_fma_per_s_1core"] = cpu_fma / dt
        out[name]["cpu_sample"] = "1 series, %d lags" % (lags + 1)
    del x, curves, tau

# ---- K9: Convergence.process on one long trajectory (host buffers in, results out) ----
rng = np.random.default_rng(0)
P, T, G = 6, 200000, 300
walk = np.cumsum(rng.integers(-1, 2, (P, T)), axis=1)
idx = (np.abs(walk) % G).astype(np.int32)
allowed = [np.linspace(0.1, 5.0, G)] * P
Cv = BC.Convergence.from_arrays(np.take(allowed[0], idx), idx, allowed, ["sigma_%d" % p for p in range(P)], outdir=None)
Cv.tau_c = np.full(P, 2.0)            # subsampling stride 5 -> 40 000 snapshots per parameter
for rngname in ("philox", "numpy"):
    np.random.seed(0)
    t0 = time.perf_counter(); Cv.process(nfold=10, nround=100, savefile=False, rng=rngname, seed=1); dt = time.perf_counter() - t0
    out["K9_process_P6_40k_snaps_nfold10_nround100_%s" % rngname] = {"s": dt, "tasks": P * 10 * 101}
np.random.seed(0)
t0 = time.perf_counter(); CO.process(idx[:1, :], Cv.tau_c[:1], [G], nfold=10, nround=10); dt = time.perf_counter() - t0
out["K9_cpu_numpy_restatement"] = {"s_scaled_to_full": dt * P * 101 / 11, "sample": "1 parameter, 10 rounds", "s": dt}
print(json.dumps(out))
