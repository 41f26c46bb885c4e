"""GPU: the CUDA Metropolis kernel (through the C ABI) against the C oracle on identical tables
and the same Philox streams -- every output must agree bit for bit."""
import numpy as np
import pytest

from helpers import cineromycin_tables, stack_lambdas, assert_outputs_equal, load_npz, tables_from_npz

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tab", "ws", "spec", "coop", "fast", "generic"], autouse=True)
def kernel(request, monkeypatch):
    """Every parity case runs through all Metropolis kernels (table-driven, warp-specialised speculative, speculative cooperative,
    cooperative lanes-per-chain, thread-per-chain specialised, generic); BCP_KERNEL is read at each launch."""
    monkeypatch.setenv("BCP_KERNEL", request.param)
    return request.param


@pytest.fixture(scope="module")
def cin():
    z, t0, t1 = cineromycin_tables()
    return stack_lambdas([t0, t1])


def run_both(tabs, n_chains, seed, n_steps, **kw):
    from biceps_b200 import engine
    from oracle import cbind
    T = engine.DeviceTables(tabs, device=0)
    ck = {k: kw[k] for k in ("chain_id0", "chain_lambda", "chain_group", "n_groups", "init_state") if k in kw}
    sk = {k: kw[k] for k in ("burn", "traj_every", "record_state_trace") if k in kw}
    got = T.chains(n_chains, seed, **ck).sample(n_steps, **sk)
    want = cbind.sample(tabs, n_chains, seed, n_steps, n_threads=8, **ck, **sk)
    return got, want


def test_cineromycin_two_lambdas_bit_exact(cin):
    n = 96
    lam = (np.arange(n) % 2).astype(np.int32)
    got, want = run_both(cin, n, 12345, 5000, burn=123, traj_every=100, record_state_trace=True, chain_lambda=lam)
    assert_outputs_equal(got, want)
    assert got["accepted"].min() > 0 and got["state_counts"].sum() == n * 5000


def test_single_chain_per_group_like_the_reference(cin):
    """One chain per histogram group == one reference sampler per chain."""
    n = 7
    got, want = run_both(cin, n, 1, 3000, traj_every=1, record_state_trace=True,
                         chain_lambda=np.ones(n, np.int32), chain_group=np.arange(n, dtype=np.int32), n_groups=n,
                         chain_id0=1000)
    assert_outputs_equal(got, want)
    assert np.array_equal(got["state_counts"].sum(axis=1), np.full(n, 3000))


def test_given_initial_states_and_ragged_chain_count(cin):
    n = 33                                     # not a multiple of the warp size
    init = (np.arange(n) * 3 % 100).astype(np.int32)
    got, want = run_both(cin, n, 77, 1000, burn=0, traj_every=7, init_state=init)
    assert_outputs_equal(got, want)


def test_state_trace_of_the_first_chains_only(cin):
    """trace_chains = k keeps the state trace of chains 0..k-1 only ([n_steps][k]); everything else unchanged."""
    from biceps_b200 import engine
    from oracle import cbind
    T = engine.DeviceTables(cin, device=0)
    lam = (np.arange(50) % 2).astype(np.int32)
    want = cbind.sample(cin, 50, 21, 3000, burn=7, traj_every=9, record_state_trace=True, chain_lambda=lam, n_threads=4)
    for k in (1, 3):
        got = T.chains(50, 21, chain_lambda=lam).sample(3000, burn=7, traj_every=9, record_state_trace=True, trace_chains=k)
        assert got["state_trace"].shape == (3000, k)
        assert np.array_equal(got["state_trace"], want["state_trace"][:, :k])
        assert_outputs_equal({x: got[x] for x in got if x != "state_trace"}, {x: want[x] for x in want if x != "state_trace"})


def test_no_snapshots_no_trace(cin):
    got, want = run_both(cin, 64, 5, 2000)
    assert_outputs_equal(got, want)
    assert got["traj_energy"].shape[0] == 0


def test_repeated_sample_calls_continue_the_chain(cin):
    from biceps_b200 import engine
    from oracle import cbind
    T = engine.DeviceTables(cin, device=0)
    blk = T.chains(40, 9)
    a = blk.sample(700, traj_every=10)
    b = blk.sample(1300, burn=50, traj_every=10)
    wa = cbind.sample(cin, 40, 9, 700, traj_every=10)
    wb = cbind.sample(cin, 40, 9, 1300, burn=50, traj_every=10, step0=700, resume=wa)
    assert_outputs_equal(a, wa)
    assert_outputs_equal(b, wb)
    one = cbind.sample(cin, 40, 9, 2050)
    assert np.array_equal(b["state"], one["state"]) and np.array_equal(b["energy"], one["energy"])


def test_three_scalar_restraints_with_reference_potentials():
    g = load_npz("apomyoglobin_tables.npz")
    tabs = tables_from_npz(g, "t_")
    got, want = run_both(tabs, 50, 31, 4000, traj_every=20, record_state_trace=True)
    assert_outputs_equal(got, want)


def test_generic_kernel_is_bit_exact_too(cin, monkeypatch):
    """The canonical layout runs the specialised kernel; BCP_FORCE_GENERIC=1 and non-canonical layouts
    (gamma restraint not last) run the generic one.  Both must match the oracle."""
    monkeypatch.setenv("BCP_FORCE_GENERIC", "1")
    lam = (np.arange(70) % 2).astype(np.int32)
    got, want = run_both(cin, 70, 4, 2500, burn=11, traj_every=13, record_state_trace=True, chain_lambda=lam)
    assert_outputs_equal(got, want)
    monkeypatch.delenv("BCP_FORCE_GENERIC")
    swapped = dict(cin)
    swapped["restraints"] = [cin["restraints"][1], cin["restraints"][0]]      # noe first: non-canonical
    got, want = run_both(swapped, 70, 4, 2500, burn=11, traj_every=13, record_state_trace=True, chain_lambda=lam)
    assert_outputs_equal(got, want)


def test_long_run_spans_several_launches(cin, kernel):
    if kernel == "generic":
        pytest.skip("slow in the generic kernel; launch splitting is host code shared by all kernels")
    """> 2^22 steps are split over kernel launches; the chain must not notice."""
    from biceps_b200 import engine
    from oracle import cbind
    T = engine.DeviceTables(cin, device=0)
    n_steps = (1 << 22) + 12345
    got = T.chains(32, 3, chain_lambda=np.ones(32, np.int32)).sample(n_steps, burn=1000, traj_every=50000)
    want = cbind.sample(cin, 32, 3, n_steps, burn=1000, traj_every=50000, chain_lambda=np.ones(32, np.int32), n_threads=16)
    assert_outputs_equal(got, want)


def test_tiny_grids_and_single_state():
    """Edge cases: 1-point gamma grid, 2-point sigma grid (wrap-around every move), a single state."""
    rng = np.random.default_rng(5)
    for S, nsig, G in [(1, 2, 1), (3, 3, 2), (5, 1, 4)]:
        sig = np.linspace(0.5, 1.5, nsig)
        def rt(kind):
            d = dict(kind=kind, name=kind, ref="uniform", ndof=3.0, sigma=sig, nlsig=3.0 * np.log(sig),
                     two_sig2=2.0 * sig ** 2, cnorm=2.75, grids=[], refsum=None)
            if kind == "gamma":
                d["grids"] = [np.linspace(0.8, 1.2, G)]
                d["sse"] = rng.uniform(1, 5, (S, G)); d["refsum"] = rng.uniform(0, 1, S); d["ref"] = "exponential"
            else:
                d["sse"] = rng.uniform(1, 5, S)
            return d
        tabs = dict(n_states=S, energy=rng.uniform(0, 2, S), logZ=0.3, restraints=[rt("scalar"), rt("gamma")])
        got, want = run_both(tabs, 40, 8, 1500, traj_every=3, record_state_trace=True)
        assert_outputs_equal(got, want)


def test_flat_gamma_rows_drift_out_of_the_prefetched_window():
    """Every gamma move is accepted (flat sse rows), so the gamma index random-walks fast: state-move
    proposals prefetched two batches ahead regularly find their 9-wide window stale, and state moves are
    accepted often (row reload path of the speculative kernel).  Short grids wrap many times."""
    rng = np.random.default_rng(11)
    S, nsig, G = 40, 7, 12
    sig = np.linspace(0.8, 1.3, nsig)
    base = rng.uniform(1, 2, S)
    gam = dict(kind="gamma", name="gamma", ref="exponential", ndof=2.0, sigma=sig, nlsig=2.0 * np.log(sig),
               two_sig2=2.0 * sig ** 2, cnorm=1.8, grids=[np.linspace(0.8, 1.2, G)],
               sse=base[:, None] + 1e-3 * rng.uniform(0, 1, (S, G)), refsum=rng.uniform(0, 1, S))
    sc = dict(kind="scalar", name="scalar", ref="uniform", ndof=2.0, sigma=sig, nlsig=2.0 * np.log(sig),
              two_sig2=2.0 * sig ** 2, cnorm=1.8, grids=[], refsum=None, sse=rng.uniform(1, 2, S))
    for rest in ([gam], [sc, gam], [sc, sc, gam], [sc, sc, sc, sc, gam]):
        tabs = dict(n_states=S, energy=rng.uniform(0, 1, S), logZ=0.1, restraints=rest)
        got, want = run_both(tabs, 70, 21, 6000, burn=17, traj_every=11, record_state_trace=True)
        assert_outputs_equal(got, want)
        assert got["sep_accepted"][:, -1].min() > 100          # state moves are accepted often


def test_neglogp_matches_oracle(cin):
    from biceps_b200 import engine
    from oracle import sampler_oracle as O
    z, t0, t1 = cineromycin_tables()
    T = engine.DeviceTables(cin, device=0)
    rng = np.random.default_rng(0)
    n = 500
    st = rng.integers(0, 100, n)
    lam = rng.integers(0, 2, n)
    idx = np.stack([rng.integers(0, g, n) for g in T.grid_len], axis=1)
    e = T.neglogp(st, idx, lam)
    want = np.array([O.neglogP(t1 if lam[i] else t0, int(st[i]), [int(x) for x in idx[i]]) for i in range(n)])
    assert np.array_equal(e, want)


def test_synthetic_c2_five_restraints_eight_lambdas():
    from biceps_b200 import synthetic
    from oracle import precompute_oracle as PO
    ens = synthetic.make_ensemble("C2", seed=1, n_states=300)
    tabs = PO.build_tables(ens["energies"], ens["lambdas"], ens["restraints"], exact_pow=False)
    n = 8 * 24
    lam = (np.arange(n) % 8).astype(np.int32)
    got, want = run_both(tabs, n, 2, 3000, burn=10, traj_every=50, chain_lambda=lam)
    assert_outputs_equal(got, want)


def test_statistical_agreement_of_independent_streams(cin, kernel):
    if kernel != "spec":
        pytest.skip("statistics do not depend on the kernel (all three are bit-identical)")
    """Posterior populations from two independent seeds agree (KL < 1e-3, north_star tolerance)."""
    from biceps_b200 import engine
    T = engine.DeviceTables(cin, device=0)
    pops = []
    for seed in (101, 202):
        o = T.chains(4096, seed, chain_lambda=np.ones(4096, np.int32)).sample(20000, burn=2000)
        c = o["state_counts"][1].astype(np.float64) + 1.0
        pops.append(c / c.sum())
    kl = float(np.sum(pops[0] * np.log(pops[0] / pops[1])))
    assert kl < 1e-3


def test_errors_are_reported_not_crashed(cin):
    from biceps_b200 import engine
    T = engine.DeviceTables(cin, device=0)
    with pytest.raises(RuntimeError):
        T.chains(4, 0, chain_lambda=np.array([0, 1, 2, 0], np.int32))     # lambda index out of range
    with pytest.raises(RuntimeError):
        T.chains(4, 0).sample(0)
    with pytest.raises(RuntimeError):
        T.neglogp([100], [[0, 0, 0]])                                     # state out of range
