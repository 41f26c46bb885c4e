"""GPU: reference-stream mode (csrc/sampler_mt.cu).  The CUDA sampler consumes numpy's MT19937 stream draw for
draw like the reference's PosteriorSampler.sample, so the trajectories the reference ITSELF produced
(tests/golden/cineromycin_mt_traj.npz, apomyoglobin_tables.npz: made by tests/golden/make_golden.py from the live
reference) are reproduced bit for bit on the GPU -- state, sigma/gamma indices, energies, histograms."""
import numpy as np
import pytest

from helpers import load_npz, tables_from_npz, cineromycin_tables

pytestmark = pytest.mark.gpu


def _run(tabs, seed, nsteps, burn):
    from biceps_b200 import engine
    T = engine.DeviceTables(tabs, device=0)
    np.random.seed(seed)
    init = np.random.randint(low=0, high=T.n_states, size=1).astype(np.int32)          # PosteriorSampler.py:29
    blk = T.chains(1, 0, init_state=init, rng_mode=1)
    o = blk.sample_numpy_stream(nsteps, burn=burn, traj_every=100, record_state_trace=True)
    after = np.random.random()
    return T, blk, o, after


@pytest.mark.parametrize("tag", ["lam1_", "lam0_"])
def test_reference_trajectory_bit_for_bit(tag):
    z, t0, t1 = cineromycin_tables()
    tabs = t1 if tag == "lam1_" else t0
    g = load_npz("cineromycin_mt_traj.npz")
    n, burn, seed = int(g[tag + "nsteps"]), int(g[tag + "burn"]), int(g[tag + "seed"])
    T, blk, o, after = _run(tabs, seed, n, burn)
    assert np.array_equal(o["traj_energy"][:, 0].view(np.int64), g[tag + "E"].view(np.int64))          # bitwise
    assert np.array_equal(o["traj_state"][:, 0], g[tag + "state"].reshape(-1))
    assert np.array_equal(o["traj_accept"][:, 0], g[tag + "accept"])
    assert np.array_equal(o["traj_indices"][:, :, 0], g[tag + "indices"])
    assert np.array_equal(o["state_trace"][:, 0], g[tag + "state_trace"].reshape(-1))
    assert np.array_equal(o["state_counts"][0] + 1, g[tag + "state_counts"])          # reference's pseudo-count (:393)
    assert np.array_equal(o["param_counts"][0], g[tag + "sampled_parameters"])
    assert int(o["accepted"][0]) == int(g[tag + "accepted"]) and int(o["total"][0]) == int(g[tag + "total"])
    assert o["energy"][0] == g[tag + "final_E"]
    assert np.array_equal(o["sep_accepted"][0] / float(o["total"][0]) * 100., g[tag + "sep_accept0"])
    # np.random was advanced by exactly what the reference's sample() would have consumed
    from oracle import philox as PX, sampler_oracle as O
    O.run_chain(tabs, PX.MTGlobalRNG(seed), n, burn=burn, traj_every=100, accept_fn=O.accept_np_exp)
    assert after == np.random.random()


def test_three_restraints_gaussian_exponential_refs():
    g = load_npz("apomyoglobin_tables.npz")
    tabs = tables_from_npz(g, "t_")
    T, blk, o, _ = _run(tabs, int(g["t_seed"]), int(g["t_nsteps"]), 0)
    assert np.array_equal(o["traj_energy"][:, 0].view(np.int64), g["t_E"].view(np.int64))
    assert np.array_equal(o["traj_indices"][:, :, 0], g["t_indices"])
    assert np.array_equal(o["state_trace"][:, 0], g["t_state_trace"].reshape(-1))
    assert np.array_equal(o["param_counts"][0], g["t_sampled_parameters"])


def test_second_call_continues_the_stream_and_errors():
    from biceps_b200 import engine
    from oracle import philox as PX, sampler_oracle as O
    z, t0, t1 = cineromycin_tables()
    T = engine.DeviceTables(t1, device=0)
    np.random.seed(21)
    init = np.random.randint(low=0, high=T.n_states, size=1).astype(np.int32)
    blk = T.chains(1, 0, init_state=init, rng_mode=1)
    a = blk.sample_numpy_stream(700, traj_every=10)
    b = blk.sample_numpy_stream(900, traj_every=10)
    rng = PX.MTGlobalRNG(21)
    r1 = O.run_chain(t1, rng, 700, traj_every=10, accept_fn=O.accept_np_exp)
    r2 = O.run_chain(t1, rng, 900, traj_every=10, accept_fn=O.accept_np_exp, chain=r1.chain)
    assert np.array_equal(a["traj_energy"][:, 0], np.array(r1.traj_E)) and np.array_equal(b["traj_energy"][:, 0], np.array(r2.traj_E))
    assert np.array_equal(b["traj_indices"][:, :, 0], np.array(r2.traj_indices))
    with pytest.raises(RuntimeError):
        T.chains(2, 0, rng_mode=1)                       # one chain only
    blk2 = T.chains(1, 0, init_state=init, rng_mode=1)
    with pytest.raises(RuntimeError):
        blk2.sample(10)                                  # no stream set
    blk2.set_mt_stream(np.arange(8, dtype=np.uint32))
    blk2.sample(100)
    with pytest.raises(RuntimeError):
        blk2.mt_consumed()                               # ran dry


def test_posterior_sampler_rng_numpy_equals_oracle_under_the_same_seed(tmp_path):
    """The drop-in API: np.random.seed(s); PosteriorSampler(ens, rng='numpy').sample(n) == the reference's run."""
    import biceps_b200 as biceps
    from oracle import philox as PX, sampler_oracle as O
    z, t0, t1 = cineromycin_tables()
    OPTS = [dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
            dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))]
    ens = biceps.Ensemble.from_arrays(1.0, z["raw_energies"], [
        dict(cls="J", model=z["obs0_model"], exp=z["obs0_exp"], restraint_index=list(z["obs0_restraint_index"]), options=OPTS[0]),
        dict(cls="noe", model=z["obs1_model"], exp=z["obs1_exp"], restraint_index=list(z["obs1_restraint_index"]), options=OPTS[1])])
    # the drop-in default (one chain, no seed argument) is the reference's stream; a seed or several chains mean Philox
    assert biceps.PosteriorSampler(ens, seed=1).rng == "philox" and biceps.PosteriorSampler(ens, nchains=4).rng == "philox"
    np.random.seed(3)
    s = biceps.PosteriorSampler(ens)
    assert s.rng == "numpy"
    s.sample(5000, burn=100)
    r = O.run_chain(s.tables_single() if hasattr(s, "tables_single") else _single(s.tables), PX.MTGlobalRNG(3), 5000,
                    burn=100, traj_every=100, accept_fn=O.accept_np_exp)
    assert [fr[3][0] for fr in s.traj.trajectory] == r.traj_state
    assert [[i for sub in fr[4] for i in sub] for fr in s.traj.trajectory] == r.traj_indices
    assert [fr[1] for fr in s.traj.trajectory] == r.traj_E
    assert s.traj.state_trace == r.state_trace
    assert np.array_equal(s.traj.state_counts, r.state_counts)


def _single(tabs):
    t = dict(tabs)
    t["energy"] = np.asarray(tabs["energy"])[0]
    t["logZ"] = float(np.asarray(tabs["logZ"])[0])
    t["n_states"] = t["energy"].shape[0]
    return t


def test_random_ensembles_on_the_mt_stream():
    """Randomised: 1-6 restraints, with / without a gamma restraint (also non-canonical positions), tiny grids and
    single-state ensembles (randint(1) draws nothing), flat tables (frequent second uniforms); the CUDA chain on
    numpy's stream against the Python oracle that calls np.random itself."""
    from biceps_b200 import engine
    from oracle import philox as PX, sampler_oracle as O
    rng = np.random.default_rng(2024)
    for case in range(40):
        R = int(rng.integers(1, 7))
        S = int(rng.choice([1, 2, 5, 40]))
        flat = bool(rng.integers(0, 2))
        scale = 1e-3 if flat else 1.0
        rest = []
        for q in range(R):
            nsig = int(rng.choice([1, 2, 3, 9, 40]))
            sig = np.linspace(0.5, 1.5, nsig) if nsig > 1 else np.array([0.9])
            ndof = float(rng.uniform(1, 6))
            d = dict(kind="scalar", name="r%d" % q, ref="uniform", ndof=ndof, sigma=sig, nlsig=ndof * np.log(sig),
                     two_sig2=2.0 * sig ** 2, cnorm=float(rng.uniform(0, 3)), grids=[], refsum=None)
            if rng.integers(0, 3) == 0:
                G = int(rng.choice([1, 2, 3, 17]))
                d.update(kind="gamma", grids=[np.linspace(0.8, 1.2, G)], sse=2.0 + scale * rng.uniform(0, 3, (S, G)))
            else:
                d["sse"] = 2.0 + scale * rng.uniform(0, 3, S)
            if rng.integers(0, 2):
                d["refsum"] = scale * rng.uniform(0, 1, S)
                d["ref"] = "exponential"
            rest.append(d)
        tabs = dict(n_states=S, energy=scale * rng.uniform(0, 2, S), logZ=float(rng.uniform(0, 1)), restraints=rest)
        steps, burn, te = int(rng.choice([1, 7, 300])), int(rng.choice([0, 3])), int(rng.choice([1, 4, 50]))
        seed = 100 + case
        T = engine.DeviceTables(tabs, device=0)
        np.random.seed(seed)
        init = np.random.randint(low=0, high=S, size=1).astype(np.int32)
        blk = T.chains(1, 0, init_state=init, rng_mode=1)
        o = blk.sample_numpy_stream(steps, burn=burn, traj_every=te, record_state_trace=True)
        after = np.random.random()
        r = O.run_chain(tabs, PX.MTGlobalRNG(seed), steps, burn=burn, traj_every=te, accept_fn=O.accept_np_exp)
        tag = "case %d R=%d S=%d" % (case, R, S)
        assert np.array_equal(o["traj_energy"][:, 0], np.array(r.traj_E)), tag
        assert np.array_equal(o["traj_indices"][:, :, 0], np.array(r.traj_indices).reshape(-1, T.n_params)), tag
        assert np.array_equal(o["state_trace"][:, 0], np.array(r.state_trace)), tag
        assert np.array_equal(o["state_counts"][0] + 1, r.state_counts), tag
        assert np.array_equal(o["param_counts"][0], np.concatenate(r.sampled_parameters)), tag
        assert int(o["accepted"][0]) == r.accepted and after == np.random.random(), tag
        blk.close(); T.close()


def test_pf_grid_reference_trajectory_bit_for_bit():
    """HDX protection factors on the 6-D nuisance grid: in this mode the PF term is summed over the residues in the
    reference's own order, so even the energies of the reference's MT-seeded run (dense 6-D tables,
    tests/golden/pf_grid.npz) are reproduced bit for bit from the raw <N_c>/<N_h> counts."""
    from helpers import pf_raw_tables
    g = load_npz("pf_grid.npz")
    raw = pf_raw_tables(g, tables_from_npz(g, "t_"))
    T, blk, o, _ = _run(raw, int(g["t_seed"]), int(g["t_nsteps"]), 0)
    assert T.n_params == 7
    assert np.array_equal(o["traj_energy"][:, 0].view(np.int64), g["t_E"].view(np.int64))
    assert np.array_equal(o["traj_indices"][:, :, 0], g["t_indices"])
    assert np.array_equal(o["traj_state"][:, 0], g["t_state"].reshape(-1))
    assert np.array_equal(o["state_trace"][:, 0], g["t_state_trace"].reshape(-1))
    assert np.array_equal(o["state_counts"][0] + 1, g["t_state_counts"])
    assert np.array_equal(o["param_counts"][0], g["t_sampled_parameters"])
    assert int(o["accepted"][0]) == int(g["t_accepted"]) and o["energy"][0] == g["t_final_E"]


def test_pf_full_default_grid_reference_trajectory_bit_for_bit():
    """The same on the reference's FULL default 6-D grid (20 x 26 x 50 x 7 x 8 x 1 = 1 456 000 cells, the grid of
    BASELINE configs[3]): tests/golden/pf_full_grid.npz is the trajectory of the live reference (dense tables of
    11.6 MB per state, MT19937 seed 11, tests/golden/make_golden_pf_full.py); here the term is evaluated on the
    fly from the raw <N_c>/<N_h> counts.  761 of the 3000 steps are accepted, so the six grid indices, sigma and
    the state all move."""
    g = load_npz("pf_full_grid.npz")
    grids = [g["grid%d" % k] for k in range(int(g["n_grids"]))]
    assert [len(x) for x in grids] == [20, 26, 50, 7, 8, 1]
    rt = dict(kind="pfgrid", name="pf", ref="exponential", ndof=float(g["ndof"]), sigma=g["sigma"], nlsig=g["nlsig"],
              two_sig2=g["two_sig2"], cnorm=float(g["cnorm"]), grids=grids, prior=None,
              pf=dict(Nc=g["Nc"], Nh=g["Nh"], exp=g["exp"], weight=g["weight"]), refsum=None, sse=None)
    raw = dict(n_states=int(g["energy"].shape[0]), energy=g["energy"], logZ=float(g["logZ"]), restraints=[rt])
    T, blk, o, _ = _run(raw, int(g["t_seed"]), int(g["t_nsteps"]), 0)
    assert T.n_params == 7
    assert np.array_equal(o["traj_energy"][:, 0].view(np.int64), g["t_E"].view(np.int64))
    assert np.array_equal(o["traj_indices"][:, :, 0], g["t_indices"])
    assert np.array_equal(o["traj_state"][:, 0], g["t_state"].reshape(-1))
    assert np.array_equal(o["state_trace"][:, 0], g["t_state_trace"].reshape(-1))
    assert np.array_equal(o["state_counts"][0] + 1, g["t_state_counts"])
    assert np.array_equal(o["param_counts"][0], g["t_sampled_parameters"])
    assert int(o["accepted"][0]) == int(g["t_accepted"]) == 761 and o["energy"][0] == g["t_final_E"]
