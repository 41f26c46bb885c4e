"""GPU: u_kn rescoring (K6) and the streaming MBAR solver (K7) against the numpy oracle and the
reference's stored tutorial results."""
import numpy as np
import pytest

from helpers import load_npz, tables_from_npz, stack_lambdas
from oracle import mbar_oracle as M

pytestmark = pytest.mark.gpu


def tutorial():
    z = load_npz("tutorial_mbar.npz")
    tabs = [tables_from_npz(z, "k%d_" % k) for k in range(2)]
    trajs = [dict(E=z["k%d_traj_E" % k], state=z["k%d_traj_state" % k], indices=z["k%d_traj_indices" % k])
             for k in range(2)]
    return z, tabs, trajs


def test_ukn_bitwise_and_tutorial_score():
    from biceps_b200 import engine, mbar
    z, tabs, trajs = tutorial()
    want = M.analysis(tabs, trajs, 100)
    T = engine.DeviceTables(stack_lambdas(tabs))
    state = np.concatenate([t["state"] for t in trajs])
    idx = np.concatenate([t["indices"] for t in trajs])
    u = mbar.ukn(T, state, idx)
    assert np.array_equal(u, want["u_kn"])                              # same fp64 op order -> same bits
    r = mbar.mbar(u, want["N_k"], state, 100)
    assert np.allclose(r["f"], want["f"], atol=1e-10)
    assert abs(r["Delta_f"][0, 1] - z["BS"][1, 0]) < 1e-8 and abs(r["dDelta_f"][0, 1] - z["BS"][1, 1]) < 1e-8
    sp = z["populations"]
    assert np.max(np.abs(sp[:, :2] - r["P"])) < 1e-12
    ok = ~np.isnan(sp[:, 2:])
    assert np.array_equal(np.isnan(sp[:, 2:]), np.isnan(r["dP"]))
    assert np.max(np.abs(sp[:, 2:][ok] - r["dP"][ok]) / sp[:, 2:][ok]) < 1e-9


@pytest.mark.parametrize("K,n", [(1, 500), (3, 1000), (11, 20000), (40, 3000)])
def test_mbar_against_oracle(K, n):
    from biceps_b200 import mbar
    rng = np.random.default_rng(K)
    kspring = np.linspace(1.0, 4.0, K)
    N_k = np.array([n + 17 * k for k in range(K)])                      # ragged sample counts
    x = np.concatenate([rng.normal(0.1 * k, 1 / np.sqrt(ks), m) for k, (ks, m) in enumerate(zip(kspring, N_k))])
    u_kn = 0.5 * kspring[:, None] * (x[None, :] - 0.1 * np.arange(K)[:, None]) ** 2
    states = (np.abs(x) * 3).astype(np.int64) % 12
    f, W, it = M.mbar_solve(u_kn, N_k)
    Df, dDf, theta = M.free_energy_differences(f, W)
    P, dP = M.populations(W, theta, states, 12)
    r = mbar.mbar(u_kn, N_k, states, 12)
    assert np.allclose(r["f"], f, atol=1e-9)
    assert np.allclose(r["theta"], theta, rtol=1e-9, atol=1e-14)
    assert np.allclose(r["dDelta_f"], dDf, rtol=1e-7, atol=1e-12)
    assert np.allclose(r["P"], P, rtol=1e-9, atol=1e-14)
    ok = ~np.isnan(dP)
    assert np.allclose(r["dP"][ok], dP[ok], rtol=1e-7)


def test_mbar_rejects_bad_input():
    from biceps_b200 import mbar
    with pytest.raises(RuntimeError):
        mbar.mbar(np.zeros((2, 10)), [5, 4])            # sum(N_k) != N
    with pytest.raises(RuntimeError):
        mbar.mbar(np.zeros((2, 10)), [5, 5], np.full(10, 3), 3)   # state out of range


def test_device_resident_score_pipeline_equals_the_host_route():
    """bcp_chains_score (snapshots stay in HBM: rescoring + MBAR) against the host route
    fetch -> concatenate per lambda in chain order -> bcp_ukn -> bcp_mbar, and against the numpy MBAR.
    Three lambdas, several seeds per lambda, chains interleaved over the lambdas."""
    from biceps_b200 import engine, mbar, synthetic
    from oracle import precompute_oracle as PO
    ens = synthetic.make_ensemble("C2", seed=4, n_states=60, n_lambda=3)
    tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
    T = engine.DeviceTables(tabs)
    n = 3 * 5 + 2                                           # ragged: lambda 0 and 1 get one more seed
    lam = (np.arange(n) % 3).astype(np.int32)
    blk = T.chains(n, 11, chain_lambda=lam)
    o = blk.sample(4000, burn=200, traj_every=10)
    got = blk.score(populations=True)
    ns = o["traj_state"].shape[0]
    order = [c for k in range(3) for c in range(n) if lam[c] == k]
    state = np.concatenate([o["traj_state"][:, c] for c in order])
    idx = np.concatenate([o["traj_indices"][:, :, c] for c in order])
    N_k = np.array([ns * int((lam == k).sum()) for k in range(3)])
    assert np.array_equal(got["N_k"], N_k)
    u = mbar.ukn(T, state, idx)
    # the stored energies are the diagonal of u_kln (Analysis.py:170) -- same bits as rescoring
    o2 = 0
    for k in range(3):
        e = np.concatenate([o["traj_energy"][:, c] for c in order if lam[c] == k])
        assert np.array_equal(u[k, o2:o2 + N_k[k]], e)
        o2 += N_k[k]
    want = mbar.mbar(u, N_k, state, 60)
    assert np.allclose(got["f"], want["f"], atol=1e-12)
    assert np.allclose(got["theta"], want["theta"], rtol=1e-10, atol=1e-15)
    assert np.allclose(got["P"], want["P"], rtol=1e-10, atol=1e-15)
    f, W, it = M.mbar_solve(u, N_k)
    assert np.allclose(got["f"], f, atol=1e-9)
    with pytest.raises(RuntimeError):
        T.chains(4, 1).score()                              # nothing sampled yet


def test_independent_runs_agree_with_the_reference_tutorial_score():
    """north_star tolerances for independent runs: |delta score| < 0.05 kT against the score the reference stored
    for its tutorial (docs/examples/Tutorials/Prep_Rest_Post_Ana/results/BS.dat: 0.3028 +- 0.0399, from its own
    MT-seeded CPU run of 2 x 200 snapshots) and KL < 1e-3 between the populations of two of our runs."""
    from biceps_b200 import engine
    z, tabs, trajs = tutorial()
    T = engine.DeviceTables(stack_lambdas(tabs))
    n = 4096
    lam = (np.arange(n) % 2).astype(np.int32)
    runs = []
    for seed in (1, 2):
        blk = T.chains(n, seed, chain_lambda=lam)
        blk.launch(20000, burn=2000, traj_every=100)
        runs.append(blk.score(populations=True))
        blk.close()
    for r in runs:
        assert abs((r["f"][1] - r["f"][0]) - z["BS"][1, 0]) < 0.05
    assert abs(runs[0]["f"][1] - runs[1]["f"][1]) < 0.05
    p, q = runs[0]["P"][:, 1], runs[1]["P"][:, 1]
    m = (p > 0) & (q > 0)
    assert float(np.sum(p[m] * np.log(p[m] / q[m]))) < 1e-3
    # the most populated states are the reference's
    assert list(np.argsort(-p)[:5]) == list(np.argsort(-z["populations"][:, 1])[:5])


def test_sample_shards_through_the_pass_entry_equal_the_single_solve():
    """The multi-GPU building block: two shards of the samples through bcp_mbar_pass_dev, their s / M summed
    (what the NCCL all-reduce does, Analysis.py:192-223 over shards) == one bcp_mbar over everything
    (f to 1e-12, theta to 1e-10); and distributed.mbar_solve over the shards converges to the same f."""
    import ctypes as C
    import torch
    from biceps_b200 import _cabi as A, distributed, mbar
    lib = A.lib()
    K, n = 11, 6000
    rng = np.random.default_rng(5)
    kspring = np.linspace(1.0, 3.0, K)
    N_k = np.array([n + 13 * k for k in range(K)], dtype=np.int64)
    x = np.concatenate([rng.normal(0.1 * k, 1 / np.sqrt(ks), m) for k, (ks, m) in enumerate(zip(kspring, N_k))])
    u_kn = np.ascontiguousarray(0.5 * kspring[:, None] * (x[None, :] - 0.1 * np.arange(K)[:, None]) ** 2)
    N = u_kn.shape[1]
    want = mbar.mbar(u_kn, N_k)
    cut = N // 3 + 7                                        # ragged shards, cutting through a run
    shards = [torch.from_numpy(np.ascontiguousarray(u_kn[:, :cut])).cuda(), torch.from_numpy(np.ascontiguousarray(u_kn[:, cut:])).cuda()]

    def pass_all(f, hess):
        s = np.zeros(K); Mm = np.zeros((K, K))
        for sh in shards:
            si = np.zeros(K); Mi = np.zeros((K, K))
            A.check(lib.bcp_mbar_pass_dev(C.c_void_p(sh.data_ptr()), sh.shape[1], K, A.i64ptr(N_k), A.dptr(A.f64(f)), int(hess), 0,
                                          C.c_void_p(0), A.dptr(si), A.dptr(Mi)), "pass")
            s += si; Mm += Mi
        return s, Mm

    s, Mm = pass_all(want["f"], True)
    assert np.allclose(s, 1.0, atol=1e-10)                  # fixed point of the MBAR equations at the single-GPU solution
    assert np.allclose(Mm, want["theta"], rtol=1e-10, atol=1e-14)
    f, theta, it = distributed.mbar_solve(pass_all, N_k, reduce=False)
    assert np.max(np.abs(f - want["f"])) < 1e-12 and it < 20
    assert np.allclose(theta, want["theta"], rtol=1e-10, atol=1e-14)


def test_score_chains_of_one_block_equals_bcp_chains_score():
    """distributed.score_chains (u_kn shard on the device + pass entry + weighted histograms, no collective for a
    single block) against bcp_chains_score on the same snapshots."""
    from biceps_b200 import distributed, engine, synthetic
    from oracle import precompute_oracle as PO
    ens = synthetic.make_ensemble("C2", seed=4, n_states=60, n_lambda=3)
    tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
    T = engine.DeviceTables(tabs)
    n = 3 * 5 + 2
    lam = (np.arange(n) % 3).astype(np.int32)
    blk = T.chains(n, 11, chain_lambda=lam)
    blk.launch(4000, burn=200, traj_every=10)
    want = blk.score(populations=True)
    got = distributed.score_chains(blk, populations=True, reduce=False)
    assert np.array_equal(got["N_k"], want["N_k"])
    assert np.max(np.abs(got["f"] - want["f"])) < 1e-12
    assert np.allclose(got["theta"], want["theta"], rtol=1e-10, atol=1e-15)
    assert np.allclose(got["P"], want["P"], rtol=1e-10, atol=1e-15)
    ok = ~np.isnan(want["dP"])
    assert np.array_equal(np.isnan(got["dP"]), np.isnan(want["dP"]))
    assert np.allclose(got["dP"][ok], want["dP"][ok], rtol=1e-8)
