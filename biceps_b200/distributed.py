"""Multi-GPU plumbing (one process per GPU, torch.distributed): chains shard across ranks with no
data-path collective; the only exchanges are the all-reduce of the per-group histograms after a
sampling pass and, for MBAR over sample shards, the all-reduce of K + K*K partial sums per
iteration.  Works with backend "nccl" (device tensors, NVLink) and, for the
CPU tests of this host logic, "gloo"."""
import ctypes as C

import numpy as np

from . import _cabi as A


def shard(n_total, rank, world):
    """Contiguous block of chains for `rank`: (first global chain id, number of chains)."""
    base, rem = divmod(int(n_total), int(world))
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


class _DevArray(object):
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_histograms(block, fused=False):
    """torch int64 views [n_groups][S] and [n_groups][param_bins] of a ChainBlock's device histograms; fused=True
    returns additionally the flat view over both (they are one block in HBM: one all-reduce sums both)."""
    import torch
    p_sc, p_pc = block.device_hist_ptrs()
    T = block.tables
    n_sc, n_pc = block.n_groups * T.n_states, block.n_groups * T.param_bins
    assert p_pc == p_sc + 8 * n_sc
    flat = torch.as_tensor(_DevArray(p_sc, (n_sc + n_pc,), "<i8"), device="cuda:%d" % T.device)
    sc = flat[:n_sc].view(block.n_groups, T.n_states)
    pc = flat[n_sc:].view(block.n_groups, T.param_bins)
    return (sc, pc, flat) if fused else (sc, pc)


def all_reduce_histograms(state_counts, param_counts, group=None):
    """Sum the per-group histograms over all ranks (torch tensors on the backend's device, or numpy
    arrays which are reduced through CPU tensors).  Returns the reduced arrays, same type as given."""
    import torch
    import torch.distributed as dist
    out = []
    for a in (state_counts, param_counts):
        if isinstance(a, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64).copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            out.append(t.numpy())
        else:
            t = a.clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            out.append(t)
    return out[0], out[1]


def mbar_solve(pass_fn, N_k, tol=1e-12, max_iter=10000, group=None, reduce=True):
    """MBAR free energies from sample shards.

    pass_fn(f, want_hess) -> (s[K], M[K][K] or None): the LOCAL sums  s_k = sum_n W_nk and
    M_kl = sum_n W_nk W_nl  over this rank's samples (on the GPU: bcp_mbar_pass_dev).  The sums are
    all-reduced, and every rank takes the same Newton / self-consistent step, so f stays replicated.
    Same iteration as csrc/mbar.cu:mbar_solve.  Returns (f, theta = W^T W, iterations)."""
    N_k = np.asarray(N_k, dtype=np.float64)
    K = N_k.shape[0]

    def global_pass(f):
        s, M = pass_fn(f, True)
        buf = np.concatenate([np.asarray(s, np.float64).ravel(), np.asarray(M, np.float64).ravel()])
        if reduce:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(buf)
            if dist.get_backend(group) == "nccl":          # NCCL reduces device tensors (K + K^2 doubles over NVLink)
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            buf = t.cpu().numpy()
        return buf[:K].copy(), buf[K:].reshape(K, K).copy()

    def gnorm(s):
        return float(np.sqrt(np.sum((N_k * (s - 1.0)) ** 2)))

    f = np.zeros(K)
    s, M = global_pass(f)
    for it in range(max_iter):
        g0 = gnorm(s)
        step = None
        if it >= 2 and K > 1:
            H = np.diag(N_k * s) - (N_k[:, None] * N_k[None, :]) * M
            try:
                d = np.linalg.solve(H[1:, 1:], -(N_k * (s - 1.0))[1:])
                cand = f.copy()
                cand[1:] += d
                s2, M2 = global_pass(cand)
                if gnorm(s2) < g0 or gnorm(s2) == 0.0:
                    step = (cand, s2, M2)
            except np.linalg.LinAlgError:
                pass
        if step is None:
            cand = np.where(N_k > 0, f - np.log(s), f)
            cand = cand - cand[0]
            s2, M2 = global_pass(cand)
            step = (cand, s2, M2)
        fn, s, M = step
        done = np.max(np.abs(fn - f)) <= tol * max(1.0, float(np.max(np.abs(fn))))
        f = fn
        if done:
            return f, M, it + 1
    raise RuntimeError("biceps_b200.distributed.mbar_solve: no convergence in %d iterations" % max_iter)


def gpu_pass_fn(d_u_kn_ptr, n_local, K, N_k_global, device=0, stream=None):
    """pass_fn for mbar_solve over a device-resident shard u_kn[K][n_local] (raw device pointer)."""
    lib = A.lib()
    N_k = np.ascontiguousarray(N_k_global, dtype=np.int64)

    def fn(f, want_hess):
        f = A.f64(f)
        s = np.zeros(K)
        M = np.zeros((K, K))
        A.check(lib.bcp_mbar_pass_dev(C.c_void_p(int(d_u_kn_ptr)), int(n_local), int(K), A.i64ptr(N_k), A.dptr(f),
                                      int(bool(want_hess)), int(device), C.c_void_p(stream or 0), A.dptr(s), A.dptr(M)),
                "bcp_mbar_pass_dev")
        return s, M
    return fn


def finish_populations(P, Q, theta):
    """P_i(l) and its 'approximate' uncertainty from the (all-reduced) weighted histograms P[K][S], Q[K][S]:
    dP = |P| sqrt(Q / P^2 + theta_ll - 2 Q / P), nan where a state was never sampled (the same
    arithmetic as csrc/mbar.cu:mbar_finish_outputs).  Returns pops[S][K], dpops[S][K]."""
    P = np.asarray(P, np.float64); Q = np.asarray(Q, np.float64)
    th = np.diag(np.asarray(theta, np.float64))[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.abs(P) * np.sqrt(np.abs(Q / (P * P) + th - 2.0 * Q / P))
    d[P == 0.0] = np.nan
    return P.T.copy(), d.T.copy()


def score_chains(block, populations=True, tol=1e-12, group=None, reduce=True):
    """BICePs scores from the snapshots of a chain block that holds a SHARD of the chains (every rank calls this
    with its own block after sampling; reduce=False: one block holds everything).  The local snapshots are rescored
    under every lambda into a device-resident u_kn shard (bcp_chains_ukn_dev), MBAR runs over the shards with one
    all-reduce of K + K^2 doubles per pass (mbar_solve + bcp_mbar_pass_dev), the weighted state histograms are
    all-reduced once (bcp_mbar_pops_dev).  Same dict as ChainBlock.score (f, theta, Delta_f, dDelta_f, P, dP, iters,
    N_k), identical on every rank."""
    import torch
    lib = A.lib()
    T = block.tables
    K, S, dev = T.n_lambda, T.n_states, T.device
    n_loc = C.c_int64(0)
    N_k = np.zeros(K, np.int64)
    A.check(lib.bcp_chains_n_samples(block._c, C.byref(n_loc), A.i64ptr(N_k)), "bcp_chains_n_samples")
    n_loc = int(n_loc.value)
    d_u = torch.empty((K, n_loc), dtype=torch.float64, device="cuda:%d" % dev)
    d_s = torch.empty(n_loc, dtype=torch.int32, device="cuda:%d" % dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    A.check(lib.bcp_chains_ukn_dev(block._c, C.c_void_p(d_u.data_ptr()), C.c_void_p(d_s.data_ptr()), C.c_void_p(st)),
            "bcp_chains_ukn_dev")

    def allred(a):
        if not reduce:
            return a
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dist.get_backend(group) == "nccl":
            t = t.cuda(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    N_k_glob = allred(N_k)
    f, theta, it = mbar_solve(gpu_pass_fn(d_u.data_ptr(), n_loc, K, N_k_glob, device=dev, stream=st), N_k_glob, tol=tol,
                              group=group, reduce=reduce)
    P = dP = None
    if populations:
        Pl = np.zeros((K, S)); Ql = np.zeros((K, S))
        A.check(lib.bcp_mbar_pops_dev(C.c_void_p(d_u.data_ptr()), n_loc, K, A.i64ptr(np.ascontiguousarray(N_k_glob, np.int64)),
                                      A.dptr(A.f64(f)), C.c_void_p(d_s.data_ptr()), S, dev, C.c_void_p(st), A.dptr(Pl),
                                      A.dptr(Ql)), "bcp_mbar_pops_dev")
        PQ = allred(np.concatenate([Pl.ravel(), Ql.ravel()]))
        P, dP = finish_populations(PQ[:K * S].reshape(K, S), PQ[K * S:].reshape(K, S), theta)
    d = np.diag(theta)
    return dict(f=f, theta=theta, Delta_f=f[None, :] - f[:, None],
                dDelta_f=np.sqrt(np.abs(d[:, None] + d[None, :] - 2.0 * theta)), P=P, dP=dP, iters=it, N_k=N_k_glob)
