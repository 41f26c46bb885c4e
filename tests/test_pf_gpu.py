"""GPU: the PF-grid Metropolis kernel (warp per chain, on-the-fly term) against the C oracle, bit for
bit, on the reference-derived golden data and on a mixed cs + pf ensemble."""
import numpy as np
import pytest

from helpers import load_npz, tables_from_npz, pf_raw_tables, assert_outputs_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tuned", "generic"], autouse=True)
def pf_kernel(request, monkeypatch):
    """Every case runs through the tuned kernel (sampler_pf2.cu: scalar restraints + one PF restraint) and the
    general one (sampler_pf.cu); BCP_PF_KERNEL is read at each launch."""
    monkeypatch.setenv("BCP_PF_KERNEL", request.param)
    return request.param


def both(tabs, n, seed, steps, **kw):
    from biceps_b200 import engine
    from oracle import cbind
    T = engine.DeviceTables(tabs)
    ck = {k: kw[k] for k in ("chain_lambda", "chain_group", "n_groups", "init_state", "chain_id0") if k in kw}
    sk = {k: kw[k] for k in ("burn", "traj_every", "record_state_trace") if k in kw}
    return T, T.chains(n, seed, **ck).sample(steps, **sk), cbind.sample(tabs, n, seed, steps, n_threads=8, **ck, **sk)


def test_pf_golden_bit_exact():
    g = load_npz("pf_grid.npz")
    raw = pf_raw_tables(g, tables_from_npz(g, "t_"))
    T, got, want = both(raw, 37, 4, 3000, burn=25, traj_every=7, record_state_trace=True)
    assert T.n_params == 7 and list(T.grid_len[1:]) == [2, 3, 2, 7, 8, 1]
    assert_outputs_equal(got, want)
    assert got["accepted"].sum() > 0


def test_pf_energies_match_dense_reference_tables():
    from biceps_b200 import engine, mbar
    from oracle import sampler_oracle as O
    g = load_npz("pf_grid.npz")
    dense = tables_from_npz(g, "t_")
    T = engine.DeviceTables(pf_raw_tables(g, dense))
    rng = np.random.default_rng(1)
    n = 200
    st = rng.integers(0, 6, n)
    idx = np.stack([rng.integers(0, gl, n) for gl in T.grid_len], axis=1)
    e = T.neglogp(st, idx)
    want = np.array([O.neglogP(dense, int(st[i]), [int(x) for x in idx[i]]) for i in range(n)])
    assert np.max(np.abs(e - want) / np.abs(want)) < 1e-12          # lane-tree vs left-to-right summation
    u = mbar.ukn(T, st, idx)
    assert u.shape == (1, n) and np.array_equal(u[0], e)


def test_mixed_cs_and_pf_two_lambdas():
    """apomyoglobin-like: three chemical-shift restraints (gaussian / exponential / uniform refs) + PF."""
    g = load_npz("pf_grid.npz")
    a = load_npz("apomyoglobin_tables.npz")
    cs = tables_from_npz(a, "t_")
    pf = pf_raw_tables(g, tables_from_npz(g, "t_"))["restraints"][0]
    S = 6
    rs = []
    for rt in cs["restraints"]:
        rt = dict(rt)
        rt["sse"] = rt["sse"][:S]
        if rt["refsum"] is not None:
            rt["refsum"] = rt["refsum"][:S]
        rs.append(rt)
    en = np.linspace(0, 2, S)
    tabs = dict(n_states=S, energy=np.array([0.0 * en, en]), logZ=np.array([np.log(S), 0.7]), restraints=rs + [pf])
    lam = (np.arange(24) % 2).astype(np.int32)
    T, got, want = both(tabs, 24, 9, 2500, burn=10, traj_every=25, chain_lambda=lam)
    assert T.n_params == 10
    assert_outputs_equal(got, want)


def test_pf_without_prior_and_uniform_reference():
    g = load_npz("pf_grid.npz")
    raw = pf_raw_tables(g, tables_from_npz(g, "t_"))
    raw["restraints"][0] = dict(raw["restraints"][0], prior=None, ref="uniform")
    T, got, want = both(raw, 8, 2, 1500, traj_every=50)
    assert_outputs_equal(got, want)
