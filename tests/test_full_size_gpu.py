"""GPU: the named full-size configurations (BASELINE configs[2] = C3: 10 000 states x 1000 observables, 8 lambdas,
4096 chains; configs[4] = C5 sizes for MBAR) through size-independent properties, plus an oracle spot check on a
subset of the chains (chains are independent and their RNG is keyed by the global chain id, so chains 0..63 of the
4096 must equal a 64-chain oracle run bit for bit)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_CHAINS, N_STEPS, K = 4096, 6000, 8


@pytest.fixture(scope="module")
def c3():
    from biceps_b200 import engine, precompute, synthetic
    ens = synthetic.make_ensemble("C3", seed=0, n_lambda=K)
    tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], device=0)   # K1/K2/K4 on the GPU
    T = engine.DeviceTables(tabs, device=0)
    lam = (np.arange(N_CHAINS) // (N_CHAINS // K)).astype(np.int32)
    return ens, tabs, T, lam


def test_c3_tables_match_the_numpy_oracle_at_full_size(c3):
    from oracle import precompute_oracle as PO
    ens, tabs, T, lam = c3
    want = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
    for a, b in zip(tabs["restraints"], want["restraints"]):
        assert np.max(np.abs(a["sse"] - b["sse"]) / np.abs(b["sse"])) < 1e-10          # tolerance stated: 1e-6
        if b.get("refsum") is not None:
            assert np.max(np.abs(a["refsum"] - b["refsum"]) / np.abs(b["refsum"])) < 1e-10
    assert np.allclose(tabs["logZ"], want["logZ"], rtol=1e-13)


def test_c3_conservation_rescoring_and_oracle_subset(c3):
    from oracle import cbind
    ens, tabs, T, lam = c3
    blk = T.chains(N_CHAINS, 11, chain_lambda=lam)
    o = blk.sample(N_STEPS, burn=500, traj_every=250, record_state_trace=False)
    per_group = (N_CHAINS // K) * N_STEPS
    # every recorded step lands in exactly one state bin and one bin of every parameter
    assert np.array_equal(o["state_counts"].sum(axis=1), np.full(K, per_group))
    off = T.hist_offsets
    for p in range(T.n_params):
        assert np.array_equal(o["param_counts"][:, off[p]:off[p] + T.grid_len[p]].sum(axis=1), np.full(K, per_group))
    assert np.all(o["total"] == N_STEPS + 500) and np.all(o["accepted"] <= o["total"]) and np.all(o["accepted"] > 0)
    # per-parameter acceptances: the parameters of one restraint move together (sigma and gamma of the NOE restraint
    # are both counted, PosteriorSampler.py:334-335), so one column per restraint + the state column add up
    first = [int(np.flatnonzero(T.rest_index == r)[0]) for r in range(int(T.rest_index.max()) + 1)]
    assert np.array_equal(o["sep_accepted"][:, first + [T.n_params]].sum(axis=1), o["accepted"])
    # idempotence: the energy a chain carries is the energy of the configuration it carries, bit for bit
    assert np.array_equal(T.neglogp(o["state"], o["indices"], lam), o["energy"])
    ns = o["traj_energy"].shape[0]
    assert ns == N_STEPS // 250
    st = o["traj_state"].reshape(-1)
    idx = o["traj_indices"].transpose(0, 2, 1).reshape(ns * N_CHAINS, T.n_params)
    assert np.array_equal(T.neglogp(st, idx, np.tile(lam, ns)), o["traj_energy"].reshape(-1))
    # chains 0..63 against the C oracle (same global chain ids, same lambda)
    w = cbind.sample(tabs, 64, 11, N_STEPS, burn=500, traj_every=250, chain_lambda=lam[:64], n_threads=8)
    for k in ("state", "indices", "energy", "accepted"):
        assert np.array_equal(o[k][:64], w[k]), k
    assert np.array_equal(o["traj_energy"][:, :64], w["traj_energy"]) and np.array_equal(o["traj_indices"][:, :, :64], w["traj_indices"])
    blk.close()


def test_c3_trajectories_do_not_depend_on_launch_block_or_kernel_split(c3, monkeypatch):
    ens, tabs, T, lam = c3
    monkeypatch.delenv("BCP_KERNEL", raising=False)
    ref = T.chains(N_CHAINS, 5, chain_lambda=lam).sample(3000, traj_every=500)
    # (i) two sample() calls instead of one, (ii) two blocks of 2048 chains with offset chain ids, (iii) the
    # thread-per-chain kernel instead of the auto-selected speculative one
    b = T.chains(N_CHAINS, 5, chain_lambda=lam)
    b.sample(1000)
    two = b.sample(2000)
    assert np.array_equal(two["state"], ref["state"]) and np.array_equal(two["energy"], ref["energy"])
    assert np.array_equal(two["state_counts"], ref["state_counts"]) and np.array_equal(two["param_counts"], ref["param_counts"])
    h = N_CHAINS // 2
    lo = T.chains(h, 5, chain_lambda=lam[:h], n_groups=K).sample(3000, traj_every=500)
    hi = T.chains(h, 5, chain_id0=h, chain_lambda=lam[h:], n_groups=K).sample(3000, traj_every=500)
    assert np.array_equal(np.concatenate([lo["energy"], hi["energy"]]), ref["energy"])
    assert np.array_equal(np.concatenate([lo["traj_indices"], hi["traj_indices"]], axis=2), ref["traj_indices"])
    assert np.array_equal(lo["state_counts"] + hi["state_counts"], ref["state_counts"])
    monkeypatch.setenv("BCP_KERNEL", "fast")
    f = T.chains(N_CHAINS, 5, chain_lambda=lam).sample(3000, traj_every=500)
    for k in ("state", "indices", "energy", "accepted", "state_counts", "param_counts", "traj_energy", "traj_indices"):
        assert np.array_equal(f[k], ref[k]), k


def test_c3_long_run_automatic_choice_is_the_table_driven_kernel_and_changes_nothing(c3, monkeypatch):
    """A long run of the named configuration goes to the table-driven kernel (trial launches of 16 384 steps, then the rest):
    same chains, histograms, counters and snapshots as the speculative kernel, bit for bit."""
    ens, tabs, T, lam = c3
    monkeypatch.delenv("BCP_KERNEL", raising=False)
    b = T.chains(N_CHAINS, 11, chain_lambda=lam)
    a = b.sample(70000, burn=1000, traj_every=1000)
    assert b.last_kernel() == "tab"
    c = b.sample(40000, traj_every=800)                       # the verdict sticks to the block: no trial, one launch
    assert b.last_kernel() == "tab"
    b.close()
    monkeypatch.setenv("BCP_KERNEL", "spec")
    b2 = T.chains(N_CHAINS, 11, chain_lambda=lam)
    a2 = b2.sample(70000, burn=1000, traj_every=1000)
    c2 = b2.sample(40000, traj_every=800)
    assert b2.last_kernel() == "spec"
    b2.close()
    for got, want in ((a, a2), (c, c2)):
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    assert a["state_counts"].sum() == N_CHAINS * 70000


def test_c2_long_run_five_restraints_on_the_table_driven_kernel_changes_nothing(monkeypatch):
    """BASELINE configs[1] (1000 states, cs_H / cs_Ca / cs_N / J / NOE, 8 lambdas x 1024 chains): five restraints leave room for a
    two-stage ring only and a quarter of the first batches are redone -- the trial launches still keep the table-driven kernel
    (rollback fraction, then the clock), and the results are those of the speculative kernel, bit for bit."""
    from biceps_b200 import engine, precompute, synthetic
    ens = synthetic.make_ensemble("C2", seed=0, n_lambda=8)
    tabs = precompute.build_tables(ens["energies"], ens["lambdas"], ens["restraints"])
    T = engine.DeviceTables(tabs)
    n = 8192
    lam = (np.arange(n) % 8).astype(np.int32)
    monkeypatch.delenv("BCP_KERNEL", raising=False)
    b = T.chains(n, 5, chain_lambda=lam)
    a = b.sample(60000, burn=500, traj_every=2000)
    assert b.last_kernel() == "tab"
    b.close()
    monkeypatch.setenv("BCP_KERNEL", "spec")
    b2 = T.chains(n, 5, chain_lambda=lam)
    a2 = b2.sample(60000, burn=500, traj_every=2000)
    assert b2.last_kernel() == "spec"
    b2.close(); T.close()
    for k in a2:
        assert np.array_equal(a[k], a2[k]), k


def test_c5_mbar_properties_at_full_size():
    """K = 11 runs x 6.4e5 samples (N = 7.04e6): f_0 = 0, gauge invariance, populations sum to one, and the
    solution satisfies the MBAR fixed-point equations to 1e-10."""
    import torch
    from biceps_b200 import mbar
    K, per, S = 11, 640000, 1000
    g = torch.Generator(device="cuda").manual_seed(3)
    lam = torch.linspace(0, 1, K, device="cuda", dtype=torch.float64)
    f_i = 3.0 * torch.randn(S, device="cuda", dtype=torch.float64, generator=g).abs()
    state = torch.cat([torch.multinomial(torch.softmax(-lam[k] * f_i, 0), per, replacement=True, generator=g) for k in range(K)])
    base = 40.0 + 2.0 * torch.randn(K * per, device="cuda", dtype=torch.float64, generator=g)
    u = (lam[:, None] * f_i[state][None, :] + base[None, :]).cpu().numpy()
    N_k = np.full(K, per, np.int64)
    st = state.cpu().numpy().astype(np.int32)
    r = mbar.mbar(u, N_k, st, S)
    assert r["f"][0] == 0.0 and np.all(np.diff(r["f"]) > 0)
    # exact answer: f_k - f_0 = -ln(Z_k / Z_0) with Z_k = sum_i exp(-lam_k f_i), within the statistical error
    exact = -(torch.logsumexp(-lam[:, None] * f_i[None, :], 1) - np.log(S)).cpu().numpy()
    assert np.max(np.abs(r["f"] - (exact - exact[0]))) < 5e-3
    assert np.allclose(r["P"].sum(axis=0), 1.0, atol=1e-9)
    # gauge: a per-sample constant and a per-run constant shift
    r2 = mbar.mbar(u + np.linspace(-5, 5, u.shape[1])[None, :], N_k)
    assert np.max(np.abs(r2["f"] - r["f"])) < 1e-9
    # fixed point: f_k = -ln sum_n exp(-u_kn - ln sum_l N_l exp(f_l - u_ln)), checked on a strided subsample of columns
    sub = u[:, ::7]
    a = r["f"][:, None] + np.log(N_k)[:, None] - u                      # full denominators are needed: compute in chunks
    den = np.zeros(u.shape[1])
    for c0 in range(0, u.shape[1], 1 << 20):
        blk = a[:, c0:c0 + (1 << 20)]
        m = blk.max(axis=0)
        den[c0:c0 + (1 << 20)] = m + np.log(np.exp(blk - m).sum(axis=0))
    fk = np.array([-(np.logaddexp.reduce(-u[k] - den)) for k in range(K)])
    assert np.max(np.abs((fk - fk[0]) - r["f"])) < 1e-9
    assert sub.shape[0] == K


def test_c4_pf_grid_properties_at_scale():
    """C4 shape (default 6-D grid 20 x 26 x 50 x 7 x 8 x 1, 107 residues) on 20 000 states (260 MB of counts, far beyond
    what the dense reference tables could hold: 232 GB): conservation, rescoring idempotence, block-split invariance
    and an oracle subset."""
    from biceps_b200 import engine, synthetic
    from oracle import cbind
    tabs = synthetic.make_pf_tables(n_states=20000, n_obs=107, seed=1)
    T = engine.DeviceTables(tabs, device=0)
    n, steps = 2048, 400
    o = T.chains(n, 3).sample(steps, burn=50, traj_every=100)
    assert int(o["state_counts"].sum()) == n * steps
    off = T.hist_offsets
    for p in range(T.n_params):
        assert int(o["param_counts"][0, off[p]:off[p] + T.grid_len[p]].sum()) == n * steps
    assert np.array_equal(T.neglogp(o["state"], o["indices"]), o["energy"])
    lo = T.chains(n // 2, 3).sample(steps, burn=50, traj_every=100)
    hi = T.chains(n // 2, 3, chain_id0=n // 2).sample(steps, burn=50, traj_every=100)
    assert np.array_equal(np.concatenate([lo["energy"], hi["energy"]]), o["energy"])
    assert np.array_equal(lo["state_counts"] + hi["state_counts"], o["state_counts"])
    w = cbind.sample(tabs, 16, 3, steps, burn=50, traj_every=100, n_threads=8)
    for k in ("state", "indices", "energy", "accepted"):
        assert np.array_equal(o[k][:16], w[k]), k
