"""CPU, world_size 2, gloo: the multi-GPU host logic -- chain sharding with globally unique Philox
streams, histogram all-reduce, and the sample-sharded MBAR solver.  The per-rank compute is done by
the oracle here (there is no GPU); what is under test is biceps_b200/distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import cineromycin_tables, stack_lambdas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from biceps_b200 import distributed as D
        from oracle import cbind, mbar_oracle as M
        z, t0, t1 = cineromycin_tables()
        tabs = stack_lambdas([t0, t1])
        n_total, n_steps = 50, 1500
        lam_all = (np.arange(n_total) % 2).astype(np.int32)
        first, n = D.shard(n_total, rank, world)
        o = cbind.sample(tabs, n, 11, n_steps, burn=100, chain_id0=first, chain_lambda=lam_all[first:first + n])
        sc, pc = D.all_reduce_histograms(o["state_counts"], o["param_counts"])
        # sample-sharded MBAR: harmonic test problem, samples split unevenly over the ranks
        rng = np.random.default_rng(0)
        K, m = 5, 3000
        ks = np.linspace(1.0, 3.0, K)
        x = np.concatenate([rng.normal(0, 1 / np.sqrt(k), m) for k in ks])
        u_kn = 0.5 * ks[:, None] * x[None, :] ** 2
        N_k = np.full(K, m)
        f0, nloc = D.shard(u_kn.shape[1] + 0, rank, world)
        cut = slice(f0, f0 + nloc) if rank == 0 else slice(f0, u_kn.shape[1])
        u_loc = u_kn[:, cut]
        logN = np.log(N_k)

        def pass_fn(f, want_hess):
            logden = M.logsumexp((f + logN)[:, None] - u_loc, axis=0)
            W = np.exp(f[:, None] - u_loc - logden[None, :])
            return W.sum(axis=1), W @ W.T
        f, theta, it = D.mbar_solve(pass_fn, N_k)
        if rank == 0:
            q.put(dict(sc=sc, pc=pc, f=f, theta=theta, it=it, first=first, n=n))
    finally:
        dist.destroy_process_group()


def test_shard_covers_all_chains():
    from biceps_b200 import distributed as D
    for n, w in [(50, 2), (4096, 8), (7, 4), (3, 8)]:
        parts = [D.shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]


@pytest.mark.timeout(300)
def test_two_ranks_gloo(oracle_lib):
    from oracle import cbind, mbar_oracle as M
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth over all 50 chains: same global chain ids -> same streams -> same histograms
    z, t0, t1 = cineromycin_tables()
    tabs = stack_lambdas([t0, t1])
    lam_all = (np.arange(50) % 2).astype(np.int32)
    one = cbind.sample(tabs, 50, 11, 1500, burn=100, chain_lambda=lam_all)
    assert np.array_equal(got["sc"], one["state_counts"]) and np.array_equal(got["pc"], one["param_counts"])
    assert got["sc"].sum() == 50 * 1500
    rng = np.random.default_rng(0)
    K, m = 5, 3000
    ks = np.linspace(1.0, 3.0, K)
    x = np.concatenate([rng.normal(0, 1 / np.sqrt(k), m) for k in ks])
    u_kn = 0.5 * ks[:, None] * x[None, :] ** 2
    f, W, it = M.mbar_solve(u_kn, np.full(K, m))
    assert np.allclose(got["f"], f, atol=1e-10)
    assert np.allclose(got["theta"], W.T @ W, rtol=1e-9)


def test_finish_populations_matches_the_oracle():
    """distributed.finish_populations (the arithmetic after the all-reduce of the weighted histograms P, Q) against
    the numpy MBAR oracle's populations and their 'approximate' uncertainties, never-sampled states included."""
    from biceps_b200 import distributed as D
    from oracle import mbar_oracle as M
    rng = np.random.default_rng(3)
    K, m, S = 4, 800, 9
    ks = np.linspace(1.0, 2.5, K)
    x = np.concatenate([rng.normal(0, 1 / np.sqrt(k), m) for k in ks])
    u_kn = 0.5 * ks[:, None] * x[None, :] ** 2
    N_k = np.full(K, m)
    states = (np.abs(x) * 3).astype(np.int64) % (S - 2)          # the last two states are never sampled
    f, W, it = M.mbar_solve(u_kn, N_k)
    Df, dDf, theta = M.free_energy_differences(f, W)
    P_want, dP_want = M.populations(W, theta, states, S)
    # what two ranks would hand to the all-reduce: P[K][S] = sum_n W_nl [s_n = i], Q = sum_n W_nl^2 [s_n = i]
    P = np.zeros((K, S)); Q = np.zeros((K, S))
    for half in (slice(0, 1500), slice(1500, None)):
        for l in range(K):
            np.add.at(P[l], states[half], W[half, l])
            np.add.at(Q[l], states[half], W[half, l] ** 2)
    pops, dpops = D.finish_populations(P, Q, theta)
    assert np.allclose(pops, P_want, rtol=1e-12, atol=1e-15)
    ok = ~np.isnan(dP_want)
    assert np.array_equal(np.isnan(dpops), np.isnan(dP_want)) and not ok[-1].any()
    assert np.allclose(dpops[ok], dP_want[ok], rtol=1e-9)
