"""Multi-GPU check (run under torchrun on N GPUs): the sample-sharded MBAR solver (bcp_mbar_pass_dev per rank + one
NCCL all-reduce of K + K^2 doubles per pass) against the single-GPU bcp_mbar on the full u_kn; also the all-reduce of
the samplers' histograms.  Prints one line per rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from biceps_b200 import distributed, mbar

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
K, per, S = 11, 64 * 2000, 500
g = torch.Generator(device="cpu").manual_seed(7)                     # the same problem on every rank
lam = torch.linspace(0, 1, K, dtype=torch.float64)
f_i = 3.0 * torch.randn(S, dtype=torch.float64, generator=g).abs()
state = torch.cat([torch.multinomial(torch.softmax(-lam[k] * f_i, 0), per, replacement=True, generator=g) for k in range(K)])
base = 40.0 + 2.0 * torch.randn(K * per, dtype=torch.float64, generator=g)
u = (lam[:, None] * f_i[state][None, :] + base[None, :]).contiguous()
N = u.shape[1]
N_k = np.full(K, per, np.int64)
lo, cnt = distributed.shard(N, rank, world)
hi = lo + cnt
u_loc = u[:, lo:hi].contiguous().cuda(local)
for _rep in range(2):                       # the second solve is warm (NCCL communicator, module load)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    f, theta, it = distributed.mbar_solve(distributed.gpu_pass_fn(u_loc.data_ptr(), hi - lo, K, N_k, device=local), N_k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
if rank == 0:
    ref = mbar.mbar(u.numpy(), N_k, device=local)
    df = np.max(np.abs(f - ref["f"]))
    dth = np.max(np.abs(theta - ref["theta"]))
    print("dist MBAR on %d GPUs: K=%d N=%d, %d iterations, %.1f ms; max |f - f_single| = %.2e, max |theta - theta_single| = %.2e"
          % (world, K, N, it, dt * 1e3, df, dth), flush=True)
    assert df < 1e-9 and dth < 1e-9
dist.barrier()
dist.destroy_process_group()
